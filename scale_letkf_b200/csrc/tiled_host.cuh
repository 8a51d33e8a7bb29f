// tiled_host.cuh -- host orchestration of the large-ensemble analysis path (tiled.cuh).  Included by
// letkf_b200.cu after the handle definition; everything runs on the handle's stream.
#pragma once

namespace {

struct TiledBufs {
  DevBuf<double> X, Ts, colsc, beta, ri, rj, lp, rz, cdiag, infl, rdiag, rloc, snorm, brk, h0, h1, misc, E, dw, U;
  DevBuf<double> bZ0, bZ1, bY0, bY1, mT, mS, E2;
  DevBuf<int> skip, ncols, cols, nobsl, idx, dims, kd, state, zsel, iters, fail, nactive, adims, solved_any;
  DevBuf<unsigned long long> snorm_bits, res, scounters;
  int *h_pinned = nullptr;   // [0..2] nactive counters of the solver loop; [1..] nobsl readback (before the solve)
  size_t h_pinned_n = 0;
  void release() {
    for (DevBuf<double> *b : {&X, &Ts, &colsc, &beta, &ri, &rj, &lp, &rz, &cdiag, &infl, &rdiag, &rloc, &snorm, &brk, &h0,
                              &h1, &misc, &E, &dw, &U, &bZ0, &bZ1, &bY0, &bY1, &mT, &mS, &E2})
      b->release();
    for (DevBuf<int> *b : {&skip, &ncols, &cols, &nobsl, &idx, &dims, &kd, &state, &zsel, &iters, &fail, &nactive, &adims,
                           &solved_any})
      b->release();
    snorm_bits.release();
    res.release();
    scounters.release();
    if (h_pinned) cudaFreeHost(h_pinned);
    h_pinned = nullptr;
    h_pinned_n = 0;
  }
};

// One batched NT GEMM launch; the tile shape follows the matrix size.
int tl_gemm(letkf_b200_handle *h, const GemmParams &P, int items) {
  if (items <= 0) return LETKF_B200_OK;
  const int mmax = P.M;
  const bool big = mmax >= 384 && (P.sym || P.N >= 128);
  if (big) {
    constexpr int BM = 128, BN = 128, ST = 3;
    const size_t smem = gemm_smem_bytes<BM, BN, ST>();
    if (!h->attr_gemm_big) {   // per handle, hence per device (the attribute is not process-wide)
      CK(cudaFuncSetAttribute(tl_gemm_kernel<BM, BN, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      h->attr_gemm_big = true;
    }
    const int tm = (mmax + BM - 1) / BM, tn = P.sym ? tm : (P.N + BN - 1) / BN;
    const int tiles = P.sym ? tm * (tm + 1) / 2 : tm * tn;
    dim3 grid((unsigned)tiles, (unsigned)(items * P.njobs));
    tl_gemm_kernel<BM, BN, ST><<<grid, (BM / 32) * (BN / 32) * 32, smem, h->stream>>>(P);
  } else {
    constexpr int BM = 64, BN = 64, ST = 3;
    const size_t smem = gemm_smem_bytes<BM, BN, ST>();
    if (!h->attr_gemm_small) {
      CK(cudaFuncSetAttribute(tl_gemm_kernel<BM, BN, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      h->attr_gemm_small = true;
    }
    const int tm = (mmax + BM - 1) / BM, tn = P.sym ? tm : (P.N + BN - 1) / BN;
    const int tiles = P.sym ? tm * (tm + 1) / 2 : tm * tn;
    dim3 grid((unsigned)tiles, (unsigned)(items * P.njobs));
    tl_gemm_kernel<BM, BN, ST><<<grid, (BM / 32) * (BN / 32) * 32, smem, h->stream>>>(P);
  }
  CK(cudaGetLastError());
  return LETKF_B200_OK;
}

// Coupled Newton-Schulz on the whole batch: on entry bY[0] holds Y0 and the per-point state is armed;
// on exit bZ[zsel[g]] (and bY[zsel[g]] when keep_y) hold the result of point g.  Without keep_y a point whose
// residual drops below 2e-3 is finished by one third- / fourth-order step (tl_step_kernel).
int tl_ns_solve(letkf_b200_handle *h, TiledBufs &T, const TiledParams &B, int nmax, bool keep_y, int *launches) {
  const int G = B.G;
  const long long sN = (long long)nmax * nmax;
  const dim3 egrid((unsigned)((sN + 1023) / 1024), (unsigned)G);
  auto base = [&](GemmParams &P) {
    std::memset(&P, 0, sizeof(P));
    P.M = P.N = P.K = nmax;
    P.mdims = B.dims;
    P.kdims = B.dims;
    P.state = B.state;
    P.sym = 1;
  };
  for (int it = 1; it <= B.max_iter + 1; ++it) {
    const int cur = it & 1, prev = cur ^ 1;
    if (it == 1) {
      tl_res_kernel<<<egrid, 256, 0, h->stream>>>(B, B.bY[0], nmax);
    } else {   // M = Z Y -> bZ[cur] (free: it holds Z of two iterations ago)
      GemmParams P;
      base(P);
      P.njobs = 1;
      P.job[0] = GemmJob{B.bZ[prev], nullptr, B.bY[prev], B.bZ[cur], sN, sN, sN, nmax, nmax, nmax};
      P.mask[0] = 1u << 0;   // points whose last step has just been completed are done
      P.res = B.res;
      int r = tl_gemm(h, P, G);
      if (r != LETKF_B200_OK) return r;
    }
    tl_step_kernel<<<1, 256, 0, h->stream>>>(B, cur, keep_y ? 0 : 1);
    CK(cudaMemcpyAsync(T.h_pinned, B.nactive, 3 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *launches += 3;
    if (T.h_pinned[0] == 0) break;
    tl_poly_kernel<<<egrid, 256, 0, h->stream>>>(B, it == 1 ? B.bY[0] : B.bZ[cur], B.mT, it == 1 ? B.bZ[1] : nullptr, nmax);
    if (T.h_pinned[1] > 0) {   // finishing steps: E^2 -> bY[cur], E^3 -> bY[prev] (Y of these points is dead), then T
      GemmParams E2;
      base(E2);
      E2.njobs = 1;
      E2.job[0] = GemmJob{B.mT, nullptr, B.mT, B.bY[cur], sN, sN, sN, nmax, nmax, nmax};
      E2.mask[0] = (1u << 3) | (1u << 4);
      int r = tl_gemm(h, E2, G);
      if (r != LETKF_B200_OK) return r;
      if (T.h_pinned[2] > 0) {
        GemmParams E3;
        base(E3);
        E3.njobs = 1;
        E3.job[0] = GemmJob{B.bY[cur], nullptr, B.mT, B.bY[prev], sN, sN, sN, nmax, nmax, nmax};
        E3.mask[0] = 1u << 4;
        r = tl_gemm(h, E3, G);
        if (r != LETKF_B200_OK) return r;
      }
      tl_finish_poly_kernel<<<egrid, 256, 0, h->stream>>>(B, B.mT, B.bY[cur], B.bY[prev], nmax);
      *launches += 3;
    }
    GemmParams P;
    base(P);
    if (it == 1) {   // Z1 = T (written by tl_poly), Y1 = T Y0
      P.njobs = 1;
      P.job[0] = GemmJob{B.mT, nullptr, B.bY[0], B.bY[1], sN, sN, sN, nmax, nmax, nmax};
      P.mask[0] = (1u << 0) | (1u << 1);
    } else {
      P.njobs = 2;
      P.job[0] = GemmJob{B.mT, nullptr, B.bZ[prev], B.bZ[cur], sN, sN, sN, nmax, nmax, nmax};
      P.job[1] = GemmJob{B.mT, nullptr, B.bY[prev], B.bY[cur], sN, sN, sN, nmax, nmax, nmax};
      P.mask[0] = (1u << 0) | (1u << 1) | (1u << 3) | (1u << 4);
      P.mask[1] = keep_y ? ((1u << 0) | (1u << 1)) : (1u << 0);
    }
    int r = tl_gemm(h, P, G);
    if (r != LETKF_B200_OK) return r;
    *launches += 2;
  }
  return LETKF_B200_OK;
}

// Analyse points [begin, end) with the tiled path, bracketed by the two events.
int launch_range_tiled(letkf_b200_handle *h, TiledBufs &T, DasParams P, long long begin, long long end, cudaEvent_t e0,
                       cudaEvent_t e1, int *launches_out) {
  const letkf_b200_config &c = h->cfg;
  const int k = c.MEMBER, n8 = round_up(k + 2, 8), kK = round_up(k, 16), maxl = h->maxl;
  const char *fe = std::getenv("LETKF_B200_TILED_FORM");
  const int form = (fe && !std::strcmp(fe, "primal")) ? 0 : (fe && !std::strcmp(fe, "dual")) ? 1 : 2;   // 2: auto
  const char *me = std::getenv("LETKF_B200_TILED_MB");
  const double budget = (me ? std::atof(me) : 4096.0) * 1048576.0;
  // When even the longest possible local list keeps every batch in the dual form, the solver matrices
  // are p x p and the batch can be much larger.
  const int pKcap = round_up(maxl, 16);
  const bool always_dual = form == 1 || (form == 2 && pKcap * 10 <= k * 6);
  const double mat = always_dual ? 6.0 * pKcap * (double)pKcap + (double)pKcap * (kK + n8) : 5.0 * n8 * (double)n8;
  const double item_bytes = mat * 8.0 + (double)maxl * 20.0 + 2.0 * kMaxNV * n8 * 8.0 + 1024.0;
  long long G = (long long)(budget / item_bytes);
  G = std::max<long long>(8, std::min<long long>(G, 4096));
  G = std::min<long long>(G, std::max<long long>(end - begin, 1));
  const size_t Gs = (size_t)G;
  int launches = 0;

  CK(T.X.ensure(Gs * kMaxNV * n8)); CK(T.Ts.ensure(Gs * n8 * kMaxNV)); CK(T.colsc.ensure(Gs * 8 * kMaxNV));
  CK(T.beta.ensure(Gs)); CK(T.ri.ensure(Gs)); CK(T.rj.ensure(Gs)); CK(T.lp.ensure(Gs)); CK(T.rz.ensure(Gs));
  CK(T.cdiag.ensure(Gs)); CK(T.infl.ensure(Gs)); CK(T.rdiag.ensure(Gs * maxl)); CK(T.rloc.ensure(Gs * maxl));
  CK(T.snorm.ensure(Gs)); CK(T.brk.ensure(Gs)); CK(T.h0.ensure(Gs)); CK(T.h1.ensure(Gs)); CK(T.misc.ensure(Gs * 4));
  CK(T.skip.ensure(Gs)); CK(T.ncols.ensure(Gs)); CK(T.cols.ensure(Gs * kMaxNV)); CK(T.nobsl.ensure(Gs));
  CK(T.idx.ensure(Gs * maxl)); CK(T.dims.ensure(Gs)); CK(T.kd.ensure(Gs)); CK(T.state.ensure(Gs)); CK(T.zsel.ensure(Gs));
  CK(T.iters.ensure(Gs)); CK(T.fail.ensure(Gs)); CK(T.nactive.ensure(4)); CK(T.adims.ensure(Gs)); CK(T.solved_any.ensure(Gs));
  CK(T.snorm_bits.ensure(Gs)); CK(T.res.ensure(Gs)); CK(T.scounters.ensure(16));
  if (T.h_pinned_n < Gs + 4) {
    if (T.h_pinned) cudaFreeHost(T.h_pinned);
    CK(cudaMallocHost((void **)&T.h_pinned, sizeof(int) * (Gs + 4)));
    T.h_pinned_n = Gs + 4;
  }
  // scratch of the search kernel (one list per resident CTA)
  const int sgrid = (int)std::max<long long>(1, std::min<long long>(G, (long long)h->num_sms * 8));
  CK(h->l_iob.ensure((size_t)sgrid * maxl)); CK(h->l_rdiag.ensure((size_t)sgrid * maxl)); CK(h->l_rloc.ensure((size_t)sgrid * maxl));
  CK(h->l_cnd.ensure((size_t)sgrid * h->ccap)); CK(h->l_cpk.ensure((size_t)sgrid * h->ccap));

  TiledParams B;
  std::memset(&B, 0, sizeof(B));
  B.n8 = n8; B.kK = kK; B.maxl = maxl; B.max_iter = P.max_sweeps + 20;
  B.X = T.X.p; B.Ts = T.Ts.p; B.colsc = T.colsc.p; B.beta = T.beta.p; B.ri = T.ri.p; B.rj = T.rj.p; B.lp = T.lp.p; B.rz = T.rz.p;
  B.skip = T.skip.p; B.ncols = T.ncols.p; B.cols = T.cols.p; B.cdiag = T.cdiag.p; B.infl = T.infl.p; B.nobsl = T.nobsl.p;
  B.idx = T.idx.p; B.rdiag = T.rdiag.p; B.rloc = T.rloc.p; B.dims = T.dims.p; B.kd = T.kd.p; B.state = T.state.p;
  B.zsel = T.zsel.p; B.iters = T.iters.p; B.fail = T.fail.p; B.snorm = T.snorm.p; B.snorm_bits = T.snorm_bits.p;
  B.brk = T.brk.p; B.h0 = T.h0.p; B.h1 = T.h1.p; B.res = T.res.p; B.misc = T.misc.p; B.nactive = T.nactive.p;
  B.adims = T.adims.p; B.solved_any = T.solved_any.p;

  CK(cudaEventRecord(e0, h->stream));
  for (long long wp0 = begin; wp0 < end; wp0 += G) {
    const int Gb = (int)std::min<long long>(G, end - wp0);
    B.G = Gb;
    B.wp0 = wp0;
    tl_load_kernel<<<Gb, 256, 0, h->stream>>>(P, B);
    ++launches;
    for (int vg = 0; vg < h->nvgroup; ++vg) {
      B.vg = vg;
      B.dual = 0;
      tl_group_kernel<<<Gb, 128, 0, h->stream>>>(P, B);
      {   // local observations of every point of the batch
        SearchParams S;
        std::memset(&S, 0, sizeof(S));
        S.T = h->d_tables.p; S.rec = h->rec.p; S.bstart = h->bstart.p;
        S.vlfac = h->vlfac_groups.p + (size_t)vg * h->nctype;
        S.ri = B.ri; S.rj = B.rj; S.lp = B.lp; S.rz = B.rz;
        S.npts = Gb; S.max_out = maxl; S.nobsl = B.nobsl; S.idx = B.idx; S.rdiag = B.rdiag; S.rloc = B.rloc;
        S.l_iob = h->l_iob.p; S.l_rdiag = h->l_rdiag.p; S.l_rloc = h->l_rloc.p; S.lcap = maxl;
        S.l_cnd = h->l_cnd.p; S.l_cpk = h->l_cpk.p; S.ccap = h->ccap;
        S.counters = T.scounters.p;
        CK(cudaMemsetAsync(T.scounters.p, 0, 16 * sizeof(unsigned long long), h->stream));
        search_kernel<<<std::min(sgrid, Gb), 128, 0, h->stream>>>(S);
      }
      tl_plan_kernel<<<(Gb + 127) / 128, 128, 0, h->stream>>>(P, B);
      CK(cudaMemcpyAsync(T.h_pinned + 1, B.nobsl, sizeof(int) * Gb, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      launches += 3;
      int pmax = 0;
      for (int g = 0; g < Gb; ++g) pmax = std::max(pmax, T.h_pinned[1 + g]);
      if (pmax > 0) {
        const int pK = round_up(pmax, 16);
        const bool dual = form == 1 || (form == 2 && pK * 10 <= k * 6);
        B.dual = dual ? 1 : 0;
        B.pK = pK;
        tl_dims_kernel<<<(Gb + 127) / 128, 128, 0, h->stream>>>(B);
        const int nmax = dual ? pK : n8;
        B.nmax = nmax;
        const size_t sN = (size_t)nmax * nmax;
        const size_t mcap = Gs * std::max((size_t)n8 * n8, sN);   // (a forced dual form may have pK > n8)
        CK(T.bZ0.ensure(mcap)); CK(T.bZ1.ensure(mcap)); CK(T.bY0.ensure(mcap)); CK(T.bY1.ensure(mcap)); CK(T.mT.ensure(mcap));
        B.bZ[0] = T.bZ0.p; B.bZ[1] = T.bZ1.p; B.bY[0] = T.bY0.p; B.bY[1] = T.bY1.p; B.mT = T.mT.p;
        const dim3 egrid((unsigned)((sN + 1023) / 1024), (unsigned)Gb);
        if (!dual) {
          CK(T.E.ensure(Gs * (size_t)n8 * pK));
          B.E = T.E.p;
          tl_gather_primal_kernel<<<dim3((unsigned)((pK + 31) / 32), (unsigned)Gb), dim3(32, 8), 0, h->stream>>>(P, B);
          GemmParams Q;   // [A | b | bd] = E E^T  -> bZ[1]
          std::memset(&Q, 0, sizeof(Q));
          Q.njobs = 1;
          Q.job[0] = GemmJob{B.E, nullptr, B.E, B.bZ[1], (long long)n8 * pK, (long long)n8 * pK, (long long)sN, pK, pK, n8};
          Q.M = Q.N = n8; Q.K = pK; Q.kdims = B.kd; Q.state = B.state; Q.mask[0] = 3u; Q.sym = 1;
          int r = tl_gemm(h, Q, Gb);
          if (r != LETKF_B200_OK) return r;
          tl_rowsum_kernel<<<dim3((unsigned)(n8 / 8), (unsigned)Gb), 256, 0, h->stream>>>(P, B);
          tl_scale_kernel<<<egrid, 256, 0, h->stream>>>(P, B, B.bZ[1], B.bY[0], nmax, 0);
          launches += 5;
          r = tl_ns_solve(h, T, B, nmax, false, &launches);
          if (r != LETKF_B200_OK) return r;
          GemmParams A;   // Ts = Z [dX | b | bd]
          std::memset(&A, 0, sizeof(A));
          A.njobs = 1;
          A.job[0] = GemmJob{B.bZ[0], B.bZ[1], B.X, B.Ts, (long long)sN, (long long)kMaxNV * n8, (long long)n8 * kMaxNV,
                             n8, n8, kMaxNV};
          A.M = n8; A.N = kMaxNV; A.K = n8; A.mdims = B.adims; A.sel = B.zsel;
          r = tl_gemm(h, A, Gb);
          if (r != LETKF_B200_OK) return r;
          ++launches;
        } else {
          CK(T.E.ensure(Gs * (size_t)pK * kK)); CK(T.dw.ensure(2 * Gs * pK)); CK(T.U.ensure(3 * Gs * pK * kMaxNV));
          CK(T.mS.ensure(Gs * sN)); CK(T.E2.ensure(Gs * (size_t)n8 * pK));
          B.E = T.E2.p;   // Yt^T in the primal layout (the apply needs both orientations)
          tl_gather_primal_kernel<<<dim3((unsigned)((pK + 31) / 32), (unsigned)Gb), dim3(32, 8), 0, h->stream>>>(P, B);
          B.E2 = T.E2.p; B.E = T.E.p; B.dw = T.dw.p; B.U = T.U.p; B.mS = T.mS.p;
          tl_gather_dual_kernel<<<dim3((unsigned)(pK / 8), (unsigned)Gb), 256, 0, h->stream>>>(P, B);
          GemmParams Q;   // S = Yt Yt^T -> mS
          std::memset(&Q, 0, sizeof(Q));
          Q.njobs = 1;
          Q.job[0] = GemmJob{B.E, nullptr, B.E, B.mS, (long long)pK * kK, (long long)pK * kK, (long long)sN, kK, kK, nmax};
          Q.M = Q.N = nmax; Q.K = kK; Q.mdims = B.dims; Q.state = B.state; Q.mask[0] = 3u; Q.sym = 1;
          int r = tl_gemm(h, Q, Gb);
          if (r != LETKF_B200_OK) return r;
          const dim3 rgrid((unsigned)(pK / 8), (unsigned)Gb);
          tl_rowsum_dual_kernel<<<rgrid, 256, 0, h->stream>>>(B, B.mS, nmax, 1);
          tl_scale_kernel<<<egrid, 256, 0, h->stream>>>(P, B, B.mS, B.bY[0], nmax, 1);
          launches += 5;
          r = tl_ns_solve(h, T, B, nmax, true, &launches);   // Z1 = (B/s1)^-1/2, Yfin = (B/s1)^1/2
          if (r != LETKF_B200_OK) return r;
          tl_dual_d_kernel<<<egrid, 256, 0, h->stream>>>(B, B.mT, nmax);          // D -> mT
          tl_dual_keep_kernel<<<egrid, 256, 0, h->stream>>>(B, nmax);             // mS <- Z1
          tl_dual_restart_kernel<<<(Gb + 127) / 128, 128, 0, h->stream>>>(P, B);
          tl_rowsum_dual_kernel<<<rgrid, 256, 0, h->stream>>>(B, B.mT, nmax, 2);
          tl_scale_kernel<<<egrid, 256, 0, h->stream>>>(P, B, B.mT, B.bY[0], nmax, 2);
          launches += 5;
          r = tl_ns_solve(h, T, B, nmax, false, &launches);  // Z2 = (D/s2)^-1/2
          if (r != LETKF_B200_OK) return r;
          {   // apply: skinny GEMMs (see tiled.cuh)
            const long long sU = (long long)kMaxNV * pK, sYt = (long long)pK * kK;
            double *U1T = B.U, *U2T = B.U + (size_t)Gs * sU, *U3T = B.U + 2 * (size_t)Gs * sU;
            GemmParams A1;   // U1T = X Yt^T
            std::memset(&A1, 0, sizeof(A1));
            A1.njobs = 1;
            A1.job[0] = GemmJob{B.X, nullptr, B.E, U1T, (long long)kMaxNV * n8, sYt, sU, n8, kK, pK};
            A1.M = kMaxNV; A1.N = pK; A1.K = round_up(k, 8);
            r = tl_gemm(h, A1, Gb);
            if (r != LETKF_B200_OK) return r;
            GemmParams A2;   // U2T = U1T Z2
            std::memset(&A2, 0, sizeof(A2));
            A2.njobs = 1;
            A2.job[0] = GemmJob{U1T, nullptr, B.bZ[0], U2T, sU, (long long)sN, sU, pK, nmax, pK};
            A2.job[0].B_alt = B.bZ[1];
            A2.selB = B.zsel;
            A2.M = kMaxNV; A2.N = pK; A2.K = pK; A2.kdims = B.dims;
            r = tl_gemm(h, A2, Gb);
            if (r != LETKF_B200_OK) return r;
            tl_dual_setb_kernel<<<dim3((unsigned)((pK + 255) / 256), (unsigned)Gb), 256, 0, h->stream>>>(B, U2T);
            GemmParams A3;   // U3T = U2T Z1
            std::memset(&A3, 0, sizeof(A3));
            A3.njobs = 1;
            A3.job[0] = GemmJob{U2T, nullptr, B.mS, U3T, sU, (long long)sN, sU, pK, nmax, pK};
            A3.M = kMaxNV; A3.N = pK; A3.K = pK; A3.kdims = B.dims;
            r = tl_gemm(h, A3, Gb);
            if (r != LETKF_B200_OK) return r;
            GemmParams A4;   // Ts = E2 U3T^T
            std::memset(&A4, 0, sizeof(A4));
            A4.njobs = 1;
            A4.job[0] = GemmJob{B.E2, nullptr, U3T, B.Ts, (long long)n8 * pK, sU, (long long)n8 * kMaxNV, pK, pK, kMaxNV};
            A4.M = n8; A4.N = kMaxNV; A4.K = pK; A4.mdims = B.adims; A4.kdims = B.dims;
            r = tl_gemm(h, A4, Gb);
            if (r != LETKF_B200_OK) return r;
            tl_dual_fin_kernel<<<dim3((unsigned)((n8 * kMaxNV + 1023) / 1024), (unsigned)Gb), 256, 0, h->stream>>>(P, B);
            launches += 6;
          }
        }
      }
      tl_update_kernel<<<Gb, 256, 0, h->stream>>>(P, B);
      ++launches;
    }
    tl_count_kernel<<<(Gb + 127) / 128, 128, 0, h->stream>>>(P, B);
    ++launches;
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(e1, h->stream));
  *launches_out += launches;
  return LETKF_B200_OK;
}

// letkf_core for ne > 128: device pointers; batches of G points through the primal tiled solve.
int core_batch_tiled(letkf_b200_handle *h, TiledBufs &T, CoreTiledParams C) {
  const int ne = C.ne, n8 = round_up(ne + 2, 8), pK = std::max(16, round_up(C.nobs, 16));
  const char *me = std::getenv("LETKF_B200_TILED_MB");
  const double budget = (me ? std::atof(me) : 4096.0) * 1048576.0;
  const double item_bytes = (5.0 * n8 * (double)n8 + (double)n8 * pK + pK + 4.0 * kMaxNV * n8) * 8.0 + 1024.0;
  long long G = (long long)(budget / item_bytes);
  G = std::max<long long>(4, std::min<long long>(G, 4096));
  G = std::min<long long>(G, std::max(C.npts, 1));
  const size_t Gs = (size_t)G, sN = (size_t)n8 * n8;
  CK(T.X.ensure(Gs * kMaxNV * n8)); CK(T.Ts.ensure(Gs * n8 * kMaxNV)); CK(T.cdiag.ensure(Gs)); CK(T.infl.ensure(Gs));
  CK(T.snorm.ensure(Gs)); CK(T.brk.ensure(Gs)); CK(T.h0.ensure(Gs)); CK(T.h1.ensure(Gs)); CK(T.misc.ensure(Gs * 4));
  CK(T.skip.ensure(Gs)); CK(T.nobsl.ensure(Gs)); CK(T.dims.ensure(Gs)); CK(T.kd.ensure(Gs)); CK(T.state.ensure(Gs));
  CK(T.zsel.ensure(Gs)); CK(T.iters.ensure(Gs)); CK(T.fail.ensure(Gs)); CK(T.nactive.ensure(4)); CK(T.adims.ensure(Gs));
  CK(T.snorm_bits.ensure(Gs)); CK(T.res.ensure(Gs)); CK(T.E.ensure(Gs * (size_t)n8 * pK)); CK(T.dw.ensure(Gs * pK));
  CK(T.bZ0.ensure(Gs * sN)); CK(T.bZ1.ensure(Gs * sN)); CK(T.bY0.ensure(Gs * sN)); CK(T.bY1.ensure(Gs * sN)); CK(T.mT.ensure(Gs * sN));
  if (T.h_pinned_n < Gs + 4) {
    if (T.h_pinned) cudaFreeHost(T.h_pinned);
    CK(cudaMallocHost((void **)&T.h_pinned, sizeof(int) * (Gs + 4)));
    T.h_pinned_n = Gs + 4;
  }
  TiledParams B;
  std::memset(&B, 0, sizeof(B));
  B.n8 = n8; B.nmax = n8; B.pK = pK; B.maxl = 0; B.max_iter = 50; B.dual = 0;
  B.X = T.X.p; B.Ts = T.Ts.p; B.cdiag = T.cdiag.p; B.infl = T.infl.p; B.snorm = T.snorm.p; B.brk = T.brk.p; B.h0 = T.h0.p;
  B.h1 = T.h1.p; B.misc = T.misc.p; B.skip = T.skip.p; B.nobsl = T.nobsl.p; B.dims = T.dims.p; B.kd = T.kd.p;
  B.state = T.state.p; B.zsel = T.zsel.p; B.iters = T.iters.p; B.fail = T.fail.p; B.nactive = T.nactive.p;
  B.adims = T.adims.p; B.snorm_bits = T.snorm_bits.p; B.res = T.res.p; B.E = T.E.p;
  B.bZ[0] = T.bZ0.p; B.bZ[1] = T.bZ1.p; B.bY[0] = T.bY0.p; B.bY[1] = T.bY1.p; B.mT = T.mT.p;
  C.sw = T.dw.p;
  DasParams P;   // only k and det are read by the shared kernels
  std::memset(&P, 0, sizeof(P));
  P.k = ne;
  P.det = C.depd ? 1 : 0;
  int launches = 0;
  const dim3 egrid((unsigned)((sN + 1023) / 1024), 1);
  for (long long p0 = 0; p0 < C.npts; p0 += G) {
    const int Gb = (int)std::min<long long>(G, C.npts - p0);
    B.G = Gb;
    C.pt0 = (int)p0;
    const dim3 eg(egrid.x, (unsigned)Gb);
    tlc_init_kernel<<<dim3((unsigned)((pK + 255) / 256), (unsigned)Gb), 256, 0, h->stream>>>(C, B);
    if (C.infl_update) tlc_rlocsum_kernel<<<Gb, 256, 0, h->stream>>>(C, B);
    tlc_gather_kernel<<<dim3((unsigned)(n8 / 8), (unsigned)Gb), 256, 0, h->stream>>>(C, B);
    GemmParams Q;   // [A | b | bd] = E E^T -> bZ[1]
    std::memset(&Q, 0, sizeof(Q));
    Q.njobs = 1;
    Q.job[0] = GemmJob{B.E, nullptr, B.E, B.bZ[1], (long long)n8 * pK, (long long)n8 * pK, (long long)sN, pK, pK, n8};
    Q.M = Q.N = n8; Q.K = pK; Q.kdims = B.kd; Q.state = B.state; Q.mask[0] = 3u; Q.sym = 1;
    int r = tl_gemm(h, Q, Gb);
    if (r != LETKF_B200_OK) return r;
    tl_rowsum_kernel<<<dim3((unsigned)(n8 / 8), (unsigned)Gb), 256, 0, h->stream>>>(P, B);
    tl_scale_kernel<<<eg, 256, 0, h->stream>>>(P, B, B.bZ[1], B.bY[0], n8, 0);
    r = tl_ns_solve(h, T, B, n8, false, &launches);
    if (r != LETKF_B200_OK) return r;
    GemmParams Z2;   // Z Z -> mT   (Pa = Z Z / s)
    std::memset(&Z2, 0, sizeof(Z2));
    Z2.njobs = 1;
    Z2.job[0] = GemmJob{B.bZ[0], B.bZ[1], B.bZ[0], B.mT, (long long)sN, (long long)sN, (long long)sN, n8, n8, n8};
    Z2.job[0].B_alt = B.bZ[1];
    Z2.sel = B.zsel; Z2.selB = B.zsel;
    Z2.M = Z2.N = Z2.K = n8; Z2.mdims = B.adims; Z2.sym = 1;
    r = tl_gemm(h, Z2, Gb);
    if (r != LETKF_B200_OK) return r;
    tlc_transm_kernel<<<Gb, 256, 0, h->stream>>>(C, B);
    tlc_out_kernel<<<dim3((unsigned)(((size_t)ne * ne + 1023) / 1024), (unsigned)Gb), 256, 0, h->stream>>>(C, B);
  }
  CK(cudaGetLastError());
  return LETKF_B200_OK;
}

}  // namespace
