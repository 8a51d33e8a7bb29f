"""Member <-> grid transposes of the ensemble state over `torch.distributed` (NCCL over
NVLink/NVSwitch on the GPUs; gloo in the CPU tests).

Twins of scale/common/common_mpi_scale.f90:

    scatter_grd_mpi_alltoall(mstart, mend, v3dg, v2dg, v3d, v2d)   :1279-1334
    gather_grd_mpi_alltoall (mstart, mend, v3d, v2d, v3dg, v2dg)   :1340-1396
    set_alltoallv_counts                                           :1401-1423
    read_ens_mpi / write_ens_mpi member loops                      :1099-1274 (the all-to-all part)

One rank per GPU plays the reference's "e-rank" (MPI_COMM_e).  Rank r holds whole members
r, r + np, r + 2 np, ... as member-major grids v3dg(nlev, nlon, nlat, nv3d); after the scatter
every rank holds its cyclically dealt columns (grd_to_buf, :1428-1440) of ALL members as
v3d(nij1, nlev, nens, nv3d), the layout das_letkf works on.  The MPI_ALLTOALL(V) of the
reference becomes one `all_to_all_single`; pack and unpack are CUDA kernels of the C ABI
(letkf_b200_grd_to_buf / buf_to_ens / ens_to_buf / buf_to_grd).

`ops` is the object providing the four pack/unpack calls and nij1_of(); by default it is the
LETKF engine (CUDA).  The CPU tests inject a host implementation so that the counts,
displacements and member bookkeeping are exercised with gloo at world_size 2.
"""
import torch
import torch.distributed as dist


def set_alltoallv_counts(mcount, ngpblock, nprocs_e, myrank_e):
    """common_mpi_scale.f90:1401-1423.  n_ens/nt_ens: counts/displacements of the blocks exchanged
    with the member-holding ranks (the first mcount); n_mem/nt_mem: blocks this rank exchanges as a
    member holder with every rank (all zero when it holds no member in this round)."""
    n_ens = [ngpblock if p < mcount else 0 for p in range(nprocs_e)]
    n_mem = [ngpblock if myrank_e < mcount else 0] * nprocs_e
    nt_ens, nt_mem = [0] * nprocs_e, [0] * nprocs_e
    for p in range(1, nprocs_e):
        nt_ens[p] = nt_ens[p - 1] + n_ens[p - 1]
        nt_mem[p] = nt_mem[p - 1] + n_mem[p - 1]
    return n_ens, nt_ens, n_mem, nt_mem


class EnsTranspose:
    def __init__(self, ops, nprocs_e, myrank_e, nlev, nv3d, nv2d, group=None, device=None, dtype=torch.float64,
                 thermo=None):
        """thermo (capi.Thermo): the member-major grids hold SCALE restart variables; state_trans is fused into the
        pack of the scatter and state_trans_inv into the unpack of the gather (common_scale.f90:1181-1280)."""
        self.ops, self.np, self.rank, self.group = ops, int(nprocs_e), int(myrank_e), group
        self.kw = {"thermo": thermo} if thermo is not None else {}
        self.nlevall = nlev * nv3d + nv2d
        self.nij1, self.nij1max = ops.nij1_of(self.np, self.rank)
        self.block = self.nij1max * self.nlevall
        self.device = device
        self.bufs = torch.zeros(self.block * self.np, dtype=dtype, device=device)
        self.bufr = torch.zeros(self.block * self.np, dtype=dtype, device=device)

    # ---- the exchange ----------------------------------------------------------------------
    def _exchange(self, mcount, to_grid_side):
        """bufs -> bufr.  to_grid_side: members (ranks < mcount) send, everyone receives
        (scatter); otherwise everyone sends, members receive (gather)."""
        if self.np == 1:
            self.bufr.copy_(self.bufs)
            return
        if mcount == self.np:   # MPI_ALLTOALL (:1309-1311)
            dist.all_to_all_single(self.bufr, self.bufs, group=self.group)
            return
        n_ens, _, n_mem, _ = set_alltoallv_counts(mcount, self.block, self.np, self.rank)
        send, recv = (n_mem, n_ens) if to_grid_side else (n_ens, n_mem)   # (:1313-1315, :1372-1374)
        dist.all_to_all_single(self.bufr[:sum(recv)], self.bufs[:sum(send)], output_split_sizes=recv,
                               input_split_sizes=send, group=self.group)

    # ---- reference entry points ------------------------------------------------------------
    def scatter_grd_mpi_alltoall(self, mstart, mend, v3dg, v2dg, v3d, v2d, nens):
        """Members mstart..mend (1-based, at most nprocs_e of them; member mstart + r lives on rank r
        as v3dg/v2dg) -> slots mstart..mend of v3d/v2d on every rank."""
        mcount = mend - mstart + 1
        assert 0 < mcount <= self.np
        if self.rank < mcount:
            self.ops.grd_to_buf(self.np, v3dg, v2dg, self.bufs, **self.kw)
        self._exchange(mcount, to_grid_side=True)
        self.ops.buf_to_ens(self.np, self.rank, nens, mstart, mend, self.bufr, v3d, v2d)

    def gather_grd_mpi_alltoall(self, mstart, mend, v3d, v2d, v3dg, v2dg, nens):
        mcount = mend - mstart + 1
        assert 0 < mcount <= self.np
        self.ops.ens_to_buf(self.np, self.rank, nens, mstart, mend, v3d, v2d, self.bufs)
        self._exchange(mcount, to_grid_side=False)
        if self.rank < mcount:
            self.ops.buf_to_grd(self.np, self.bufr, v3dg, v2dg, **self.kw)

    # ---- member loops of read_ens_mpi / write_ens_mpi (:1099-1274) -----------------------------
    def rounds(self, nmem):
        """(it, im, mstart, mend): round `it` moves members mstart..mend, this rank holds member im
        (None when it has none in this round)."""
        nit = (nmem + self.np - 1) // self.np
        for it in range(nit):
            im = self.rank + 1 + it * self.np
            mstart = 1 + it * self.np
            mend = min((it + 1) * self.np, nmem)
            yield it, (im if im <= nmem else None), mstart, mend

    def read_ens(self, my_members3d, my_members2d, v3d, v2d, nmem, nens):
        """my_members3d[it]: the v3dg of the member this rank holds in round it (or None)."""
        for it, im, mstart, mend in self.rounds(nmem):
            g3 = my_members3d[it] if im is not None else None
            g2 = my_members2d[it] if (im is not None and my_members2d is not None) else None
            self.scatter_grd_mpi_alltoall(mstart, mend, g3, g2, v3d, v2d, nens)

    def write_ens(self, v3d, v2d, my_members3d, my_members2d, nmem, nens):
        for it, im, mstart, mend in self.rounds(nmem):
            g3 = my_members3d[it] if im is not None else None
            g2 = my_members2d[it] if (im is not None and my_members2d is not None) else None
            self.gather_grd_mpi_alltoall(mstart, mend, v3d, v2d, g3, g2, nens)


class EnsTransposeP2P:
    """read_ens_mpi / write_ens_mpi twins on the ONE-PASS transposes (letkf_b200_scatter_grd_p2p / _gather_grd_p2p): every
    rank reads its own arrays and writes straight into the receiving rank's array over NVLink, no pack buffer, no
    collective call.  Ranks of one node; peer arrays are mapped once with CUDA IPC (handles exchanged through
    torch.distributed.all_gather_object).  Barriers around the calls play the role of the blocking MPI_ALLTOALL."""

    def __init__(self, eng, nprocs_e, myrank_e, group=None, thermo=None):
        self.eng, self.np, self.rank, self.group, self.thermo = eng, int(nprocs_e), int(myrank_e), group, thermo
        self._maps = {}

    def _addresses(self, key, owner, tensor):
        """device addresses of `tensor` (same role on every rank) on all ranks, as seen from this GPU; None entries allowed.

        The handle exchange is a COLLECTIVE, so whether it runs must be decided identically on every rank: the cache is
        keyed by `key` and validated by the identity of `owner` (the Python object the caller passed: the same object on
        every rank at the same point of the program), never by this rank's pointer alone -- a rank that holds no member
        in a round (tensor None) would otherwise skip an exchange its peers enter (hang at 4 / 8 ranks with 50 members)."""
        ptr = None if tensor is None else tensor.data_ptr()
        hit = self._maps.get(key)
        if hit is not None and hit[0] is owner:
            if hit[1] != ptr:
                raise RuntimeError("EnsTransposeP2P: tensor %r of a registered owner was re-allocated; pass a new owner "
                                   "object (a new list / tensor) on every rank instead" % (key,))
            return hit[2]
        if self.np == 1:
            addrs = [ptr]
        else:
            desc = None if tensor is None else self.eng.peer_export(tensor)
            allg = [None] * self.np
            dist.all_gather_object(allg, desc, group=self.group)
            addrs = [ptr if r == self.rank else (None if d is None else self.eng.peer_open(d)) for r, d in enumerate(allg)]
        self._maps[key] = (owner, ptr, addrs)   # the strong reference to `owner` keeps its id from being reused
        return addrs

    def _barrier(self):
        torch.cuda.synchronize()
        if self.np > 1:
            dist.barrier(group=self.group)

    def rounds(self, nmem):
        nit = (nmem + self.np - 1) // self.np
        for it in range(nit):
            im = self.rank + 1 + it * self.np
            yield it, (im if im <= nmem else None), 1 + it * self.np, min((it + 1) * self.np, nmem)

    def read_ens(self, my_members3d, v3d, nmem, nens):
        """my_members3d[it]: member-major grid of the member this rank holds in round it (or None) -> v3d on every rank"""
        peers = self._addresses("v3d", v3d, v3d)
        self._barrier()                       # nobody still reads the previous contents of v3d
        for it, im, mstart, mend in self.rounds(nmem):
            if im is not None:
                self.eng.scatter_grd_p2p(self.np, self.rank, nens, im, my_members3d[it], None, peers, thermo=self.thermo)
        self._barrier()

    def write_ens(self, v3d, my_members3d, nmem, nens):
        """v3d (e.g. the analysis) on every rank -> member-major grids my_members3d[it] on the member-holding ranks"""
        grids = [self._addresses(("g", it), my_members3d, my_members3d[it]) for it, _, _, _ in self.rounds(nmem)]
        self._barrier()
        for it, im, mstart, mend in self.rounds(nmem):
            self.eng.gather_grd_p2p(self.np, self.rank, nens, mstart, mend, v3d, None, grids[it][:mend - mstart + 1],
                                    thermo=self.thermo)
        self._barrier()
