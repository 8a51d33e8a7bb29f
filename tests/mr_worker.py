"""Worker of the world_size-2 gloo tests: the member<->grid transposes of scale_letkf_b200.transpose
with HOST pack/unpack (the oracle's restatement of grd_to_buf / buf_to_grd, test infrastructure)
standing in for the CUDA kernels, so that counts, displacements and member bookkeeping of the
N > 1 path run on CPU."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class HostOps:
    """grd_to_buf / buf_to_ens / ens_to_buf / buf_to_grd on CPU torch tensors via oracle/."""

    def __init__(self, nlon, nlat, nlev, nv3d, nv2d):
        from oracle import oracle_py
        oracle_py.build()
        self.o = oracle_py
        self.L = oracle_py.lib()
        self.d = (nlon, nlat, nlev, nv3d, nv2d)

    @staticmethod
    def _p(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def nij1_of(self, np_, rank):
        return self.o.nij1(self.d[0], self.d[1], np_, rank)

    def _trans(self, v3dg, thermo, inverse):
        nlon, nlat, nlev, nv3d, _ = self.d
        self.L.oracle_state_trans(C.byref(thermo), int(inverse), nlev, nlon, nlat, nv3d, 6, self._p(v3dg))

    def grd_to_buf(self, np_, v3dg, v2dg, bufs, thermo=None):
        if thermo is not None:   # host stand-in of the fused pack: state_trans on a copy, then pack
            v3dg = v3dg.clone()
            self._trans(v3dg, thermo, False)
        self.L.oracle_grd_to_buf(*self.d, np_, self._p(v3dg), self._p(v2dg), self._p(bufs))

    def buf_to_grd(self, np_, bufr, v3dg, v2dg, thermo=None):
        self.L.oracle_buf_to_grd(*self.d, np_, self._p(bufr), self._p(v3dg), self._p(v2dg))
        if thermo is not None:
            self._trans(v3dg, thermo, True)

    def buf_to_ens(self, np_, rank, nens, mstart, mend, bufr, v3d, v2d):
        self.L.oracle_buf_to_ens(*self.d, np_, rank, nens, mstart, mend, self._p(bufr), self._p(v3d), self._p(v2d))

    def ens_to_buf(self, np_, rank, nens, mstart, mend, v3d, v2d, bufs):
        self.L.oracle_ens_to_buf(*self.d, np_, rank, nens, mstart, mend, self._p(v3d), self._p(v2d), self._p(bufs))


def member_grid(m, nlon, nlat, nlev, nv3d, nv2d):
    """deterministic member-major fields: value encodes (member, var, j, i, k)"""
    k, i, j, n = np.meshgrid(np.arange(nlev), np.arange(nlon), np.arange(nlat), np.arange(nv3d), indexing="ij")
    v3 = (m * 1e6 + n * 1e5 + j * 1e3 + i * 10 + k).astype(np.float64)      # (nlev, nlon, nlat, nv3d)
    i2, j2, n2 = np.meshgrid(np.arange(nlon), np.arange(nlat), np.arange(nv2d), indexing="ij")
    v2 = (-(m * 1e6 + n2 * 1e5 + j2 * 1e3 + i2 * 10)).astype(np.float64)    # (nlon, nlat, nv2d)
    return v3, v2


def run(rank, world, port, nmem, q):
    import torch
    import torch.distributed as dist
    from scale_letkf_b200 import synth
    from scale_letkf_b200.transpose import EnsTranspose, set_alltoallv_counts
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nlon, nlat, nlev, nv3d, nv2d = 7, 5, 3, 2, 1
        nens = nmem + 1
        ops = HostOps(nlon, nlat, nlev, nv3d, nv2d)
        tr = EnsTranspose(ops, world, rank, nlev, nv3d, nv2d)
        nij1 = tr.nij1
        # Fortran-ordered arrays as flat torch tensors (memory order = Fortran order)
        F = lambda a: torch.from_numpy(np.ascontiguousarray(a.ravel(order="F")))
        mine3, mine2, ids = [], [], []
        for it, im, mstart, mend in tr.rounds(nmem):
            if im is None:
                mine3.append(None); mine2.append(None); ids.append(None)
            else:
                v3, v2 = member_grid(im, nlon, nlat, nlev, nv3d, nv2d)
                mine3.append(F(v3)); mine2.append(F(v2)); ids.append(im)
        v3d = torch.zeros(nij1 * nlev * nens * nv3d, dtype=torch.float64)
        v2d = torch.zeros(nij1 * nens * nv2d, dtype=torch.float64)
        tr.read_ens(mine3, mine2, v3d, v2d, nmem, nens)
        a3 = v3d.numpy().reshape((nij1, nlev, nens, nv3d), order="F")
        a2 = v2d.numpy().reshape((nij1, nens, nv2d), order="F")
        ilon, ilat = synth.column_deal(nlon, nlat, world, rank)
        for m in range(1, nmem + 1):
            g3, g2 = member_grid(m, nlon, nlat, nlev, nv3d, nv2d)
            assert np.array_equal(a3[:, :, m - 1, :], g3[:, ilon - 1, ilat - 1, :].transpose(1, 0, 2)), ("scatter3", m)
            assert np.array_equal(a2[:, m - 1, :], g2[ilon - 1, ilat - 1, :]), ("scatter2", m)
        # the way back
        out3 = [None if t is None else torch.zeros_like(t) for t in mine3]
        out2 = [None if t is None else torch.zeros_like(t) for t in mine2]
        tr.write_ens(v3d, v2d, out3, out2, nmem, nens)
        for t, o in list(zip(mine3, out3)) + list(zip(mine2, out2)):
            if t is not None:
                assert torch.equal(t, o), "gather"
        # set_alltoallv_counts as the reference computes it
        n_ens, nt_ens, n_mem, nt_mem = set_alltoallv_counts(1, 10, world, rank)
        assert n_ens == [10] + [0] * (world - 1) and nt_ens[1] == 10
        assert n_mem == ([10] * world if rank == 0 else [0] * world)
        q.put((rank, "ok"))
    except Exception as e:   # report instead of hanging the peer
        import traceback
        q.put((rank, "FAIL " + repr(e) + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def run_thermo(rank, world, port, q):
    """world_size-2 gloo: EnsTranspose(thermo=...) -- state_trans fused into the scatter's pack, state_trans_inv into
    the gather's unpack (host stand-ins): the dealt columns hold u, v, w, T, p and the way back restores the restart
    variables (common_scale.f90:1181-1280 inside common_mpi_scale.f90:1099-1274)."""
    import torch
    import torch.distributed as dist
    from scale_letkf_b200 import synth, capi
    from scale_letkf_b200.transpose import EnsTranspose
    from oracle import oracle_py
    from test_state_trans import restart_state, thermo
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nlon, nlat, nlev, nv3d, nmem = 7, 5, 3, 11, 3
        nens = nmem + 1
        t = thermo()
        ops = HostOps(nlon, nlat, nlev, nv3d, 0)
        tr = EnsTranspose(ops, world, rank, nlev, nv3d, 0, thermo=t)
        F = lambda a: torch.from_numpy(np.ascontiguousarray(a.ravel(order="F")))
        mine = []
        for it, im, mstart, mend in tr.rounds(nmem):
            mine.append(None if im is None else F(restart_state(nlev, nlon, nlat, seed=200 + im)))
        v3d = torch.zeros(tr.nij1 * nlev * nens * nv3d, dtype=torch.float64)
        tr.read_ens(mine, None, v3d, None, nmem, nens)
        a3 = v3d.numpy().reshape((tr.nij1, nlev, nens, nv3d), order="F")
        ilon, ilat = synth.column_deal(nlon, nlat, world, rank)
        for m in range(1, nmem + 1):
            ref = oracle_py.state_trans(t, restart_state(nlev, nlon, nlat, seed=200 + m), inverse=False)
            assert np.array_equal(a3[:, :, m - 1, :], ref[:, ilon - 1, ilat - 1, :].transpose(1, 0, 2)), ("scatter+trans", m)
        out = [None if x is None else torch.zeros_like(x) for x in mine]
        tr.write_ens(v3d, None, out, None, nmem, nens)
        for x, o in zip(mine, out):
            if x is not None:
                err = float((o - x).abs().max() / x.abs().max())
                assert err <= 1e-12, ("gather+trans_inv", err)
        q.put((rank, "ok"))
    except Exception as e:
        import traceback
        q.put((rank, "FAIL " + repr(e) + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def run_gpu(rank, world, port, q, k=6):
    """One rank per GPU, NCCL: members -> (CUDA pack, all-to-all, CUDA unpack) -> ensmean_grd ->
    das_letkf on this rank's columns -> the way back; checked against the single-domain oracle."""
    import torch
    import torch.distributed as dist
    import scale_letkf_b200 as sl
    from scale_letkf_b200 import synth
    from scale_letkf_b200.transpose import EnsTranspose
    from oracle import oracle_py
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        cfg = synth.config_c2(nlon=12, nlat=12, nlev=4, member=k)
        for t in range(24):
            cfg.HORI_LOCAL[t] = 60.0e3
        nens, nv3d, nlev, nlon, nlat = k + 1, cfg.nv3d, cfg.nlev, cfg.nlon, cfg.nlat
        obs = synth.make_sonde_obs(cfg, 10, 30, nlevobs=6, seed_no=51)
        rig, rjg, hgt = synth.make_grid(cfg)                                   # whole plane, deal order of np = 1
        full = synth.make_state(cfg, rig, rjg, hgt, seed_no=52)                 # (nij, nlev, nens, nv3d)
        full[:, :, k, :] = 0.0
        grids = [np.asfortranarray(full[:, :, m, :].reshape((nlon, nlat, nlev, nv3d), order="F").transpose(2, 0, 1, 3))
                 for m in range(k)]                                             # v3dg(nlev, nlon, nlat, nv3d) per member
        rig1, rjg1, hgt1 = synth.make_grid(cfg, nprocs_e=world, myrank_e=rank)
        nij1 = len(rig1)
        eng = sl.LETKF(cfg, device=rank)
        eng.set_letkf_obs(obs)
        eng.set_common_mpi_grid(rig1, rjg1, hgt1)
        tr = EnsTranspose(eng, world, rank, nlev, nv3d, 0, device=dev)
        assert tr.nij1 == nij1
        F = lambda a: torch.from_numpy(np.ascontiguousarray(a.ravel(order="F"))).to(dev)
        mine = [None if im is None else F(grids[im - 1]) for _, im, _, _ in tr.rounds(k)]
        v3d = torch.zeros((nv3d, nens, nlev, nij1), dtype=torch.float64, device=dev)   # Fortran (nij1,nlev,nens,nv3d)
        tr.read_ens(mine, None, v3d, None, k, nens)
        # the one-pass transposes over peer memory (CUDA IPC + NVLink stores) must give the same bits
        from scale_letkf_b200.transpose import EnsTransposeP2P
        p2p = EnsTransposeP2P(eng, world, rank)
        v3d_p = torch.full_like(v3d, -7.0)
        p2p.read_ens(mine, v3d_p, k, nens)
        assert torch.equal(v3d_p[:, :k], v3d[:, :k]), "p2p scatter differs from the NCCL path"
        eng.ensmean_grd(v3d)
        # the same columns cut out of the whole-plane state
        ilon, ilat = synth.column_deal(nlon, nlat, world, rank)
        cols = (ilon - 1) + (ilat - 1) * nlon
        ref_in = np.asfortranarray(full[cols])
        oracle_py.ensmean_grd(k, ref_in)
        got_in = v3d.cpu().numpy().transpose(3, 2, 1, 0)
        assert np.array_equal(got_in, ref_in), "transposed state differs from the column cut"
        o = oracle_py.Oracle(cfg)
        o.set_obs(obs)
        o.set_grid(rig1, rjg1, hgt1)
        ref = o.das_letkf(ref_in.copy(order="F"))
        out = eng.das_letkf(v3d)
        anal = out["anal3d"]
        a = anal.cpu().numpy().transpose(3, 2, 1, 0)[:, :, :k, :]
        b = ref["anal3d"][:, :, :k, :]
        sc = np.maximum(np.abs(b).max(axis=(0, 1, 2), keepdims=True), 1e-300)
        err = float((np.abs(a - b) / sc).max())
        assert err <= 1e-10, f"analysis mismatch {err:.3e}"
        # way back: analysis members as member-major grids on their owner ranks
        outg = [None if t is None else torch.zeros_like(t) for t in mine]
        tr.write_ens(anal, None, outg, None, k, nens)
        outp = [None if t is None else torch.full_like(t, -7.0) for t in mine]
        p2p.write_ens(anal, outp, k, nens)
        for a_, b_ in zip(outp, outg):
            assert (a_ is None and b_ is None) or torch.equal(a_, b_), "p2p gather differs from the NCCL path"
        # a second set of output grids (what bench.py's cycle leg does): the handle exchange runs again on EVERY rank,
        # also on one that holds no member in the last round (k odd)
        outp2 = [None if t is None else torch.full_like(t, -9.0) for t in mine]
        p2p.write_ens(anal, outp2, k, nens)
        p2p.write_ens(anal, outp, k, nens)
        for a_, b_ in zip(outp2, outg):
            assert (a_ is None and b_ is None) or torch.equal(a_, b_), "p2p gather into a second set of grids"
        # every rank gathers all analysis columns through the oracle to check its own members
        allb = [None] * world
        dist.all_gather_object(allb, (cols, b))
        whole = np.zeros((nlon * nlat, nlev, k, nv3d))
        for c, bb in allb:
            whole[c] = bb
        for (_, im, _, _), g in zip(tr.rounds(k), outg):
            if im is None:
                continue
            want = whole[:, :, im - 1, :].reshape((nlon, nlat, nlev, nv3d), order="F").transpose(2, 0, 1, 3)
            gotg = g.cpu().numpy().reshape((nlev, nlon, nlat, nv3d), order="F")
            sc = np.maximum(np.abs(want).max(axis=(0, 1, 2), keepdims=True), 1e-300)
            assert float((np.abs(gotg - want) / sc).max()) <= 1e-10, "gathered analysis member"
        eng.close()
        q.put((rank, "ok"))
    except Exception as e:
        import traceback
        q.put((rank, "FAIL " + repr(e) + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


class _StubP2PEngine:
    """Stand-in of the engine for EnsTransposeP2P on CPU: 'exports' a tensor as (rank, data_ptr), 'opens' it as an
    integer, counts the calls.  The p2p kernels themselves need CUDA IPC (covered by the 2-GPU test)."""

    def __init__(self, rank):
        self.rank, self.exports, self.calls = rank, 0, []

    def peer_export(self, t):
        self.exports += 1
        return (self.rank, t.data_ptr())

    def peer_open(self, d):
        return 1000 + d[0]

    def scatter_grd_p2p(self, np_, rank, nens, im, g3, g2, peers, thermo=None):
        assert g3 is not None and len(peers) == np_ and all(p is not None for p in peers)
        self.calls.append(("s", im))

    def gather_grd_p2p(self, np_, rank, nens, mstart, mend, v3d, v2d, peers, thermo=None):
        assert len(peers) == mend - mstart + 1 and all(p is not None for p in peers), peers
        self.calls.append(("g", mstart, mend))


def run_p2p_bookkeeping(rank, world, port, nmem, q):
    """EnsTransposeP2P's handle exchange is a collective: every rank must enter it the same number of times, also when
    the rank holds no member in the last round (its grid is None) and the caller switches between two grid lists --
    the sequence bench.py's cycle leg produces (write gin, read gin, write gout, ...).  A mismatch hangs: the test
    harness times out."""
    try:
        import datetime
        import torch
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=60))
        from scale_letkf_b200.transpose import EnsTransposeP2P
        orig = dist.all_gather_object
        ncoll = [0]

        def counting(*a, **kw):
            ncoll[0] += 1
            return orig(*a, **kw)
        dist.all_gather_object = counting
        eng = _StubP2PEngine(rank)
        tr = EnsTransposeP2P(eng, world, rank)
        torch.cuda.synchronize = lambda *a, **kw: None      # CPU-only process
        rounds = list(tr.rounds(nmem))
        mk = lambda: [torch.zeros(4, dtype=torch.float64) if im is not None else None for _, im, _, _ in rounds]
        gin, gout = mk(), mk()
        v3d, anal = torch.zeros(8, dtype=torch.float64), torch.zeros(8, dtype=torch.float64)
        tr.write_ens(v3d, gin, nmem, nmem + 1)
        for _ in range(3):
            tr.read_ens(gin, v3d, nmem, nmem + 1)
            tr.write_ens(anal, gout, nmem, nmem + 1)
        tr.read_ens(gin, v3d, nmem, nmem + 1)
        tr.write_ens(v3d, gout, nmem, nmem + 1)
        n = torch.tensor([ncoll[0]], dtype=torch.int64)
        lo, hi = n.clone(), n.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert int(lo) == int(hi) == 1 + 2 * len(rounds), (int(lo), int(hi), len(rounds))   # v3d, gin[it], gout[it]: once each
        mine = sum(1 for _, im, _, _ in rounds if im is not None)
        assert sum(1 for c in eng.calls if c[0] == "s") == 4 * mine
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:   # pragma: no cover
        import traceback
        q.put((rank, "FAIL " + repr(e) + traceback.format_exc()))
