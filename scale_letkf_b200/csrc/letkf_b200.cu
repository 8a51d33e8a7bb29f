// letkf_b200.cu -- C ABI of the B200-native LETKF analysis path (include/letkf_b200.h).
// Host side: handle, configuration, observation tables, kernel launches.  No CPU compute
// path exists: every entry point that produces numbers launches sm_100a kernels.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include <cuda.h>

#include "../../include/letkf_b200.h"
#include "aux_kernels.cuh"
#include "das_kernel.cuh"
#include "das_ns_kernel.cuh"
#include "tiled.cuh"
#include "radar.cuh"

using namespace letkf;

namespace {

const int ELEM_UID[LETKF_B200_NID_OBS] = {2819, 2820, 3073, 3074, 3330, 3331, 14593, 19999,
                                          4001, 4004, 4002, 4003, 8800, 99991, 99992, 99993};
const int ID_PS = 14593, ID_RAIN = 19999, ID_REF = 4001, ID_RE0 = 4004, ID_VR = 4002;

int uid_obs(int elm) {   // common_obs_scale.f90:171-211
  for (int i = 0; i < LETKF_B200_NID_OBS; ++i)
    if (ELEM_UID[i] == elm) return i + 1;
  return -1;
}
int uid_obs_varlocal(int elm) {   // common_obs_scale.f90:216-242
  switch (elm) {
    case 2819: case 2820: return 1;
    case 3073: case 3074: return 2;
    case 3330: case 3331: return 3;
    case 14593: return 4;
    case 19999: return 5;
    case 99991: case 99992: case 99993: return 6;
    case 4001: case 4004: case 4003: return 7;
    case 4002: return 8;
    case 8800: return 9;
    default: return -1;
  }
}

// Size class of the tensor-core solver (das_ns_kernel<NB>): NB odd 8-row blocks with k + 2 <= 8 NB
// (two padding columns carry dep / depd); 0 = ensemble too large, Cholesky + Jacobi path.
int ns_class(int k) {
  if (k + 2 <= 24) return 3;
  if (k + 2 <= 40) return 5;
  if (k + 2 <= 56) return 7;
  if (k + 2 <= 72) return 9;
  if (k + 2 <= 104) return 13;
  return 0;
}
// doubles per row of the sorted obs table: the padded width of the solver class, so that a row is
// copied to shared memory as is
int ns_row_doubles(int k) {
  const int nb = ns_class(k);
  return nb ? 8 * nb : round_up(k + 2, 8);   // tiled path: n8
}

template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;              // owns its allocation: function-local buffers are freed on every return path
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  cudaError_t ensure(size_t count) {
    if (count <= n && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    cudaError_t e = cudaMalloc((void **)&p, std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

struct TiledBufs;   // scratch of the large-ensemble path (tiled_host.cuh), allocated on first use

}  // namespace

struct letkf_b200_handle {
  letkf_b200_config cfg;
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  // grid
  int nij1 = 0;
  DevBuf<double> rig1, rjg1, hgt1;
  // observations
  bool obs_set = false;
  int nobstotal = 0, nctype = 0, nensobs = 0, ldens = 0, nbuckets = 0;
  SearchTables tables;
  std::vector<letkf_b200_ctype_info> ctinfo;
  std::vector<int> h_bstart, h_s2o;
  std::vector<double> h_logp;   // host-computed ln(mean pressure) of the last host-buffer das call
  bool radar_only = true;
  int maxl = 1;
  DevBuf<SearchTables> d_tables;
  DevBuf<ObsRec> rec;
  DevBuf<int> bstart, s2o;
  DevBuf<double> sval, sens;
  // variable localisation groups
  int nvgroup = 1;
  int vgroup[kMaxNV], vfirst[kMaxNV];
  std::vector<double> h_vlfac;   // [nvar][nctype] factors for every variable (obs_local twin)
  DevBuf<double> vlfac_groups, vlfac_one;
  // scratch
  DevBuf<int> l_iob;
  DevBuf<double> l_rdiag, l_rloc, l_cnd;
  DevBuf<unsigned> l_cpk;
  DevBuf<double> m0_scratch;   // das_ns_kernel: A / s of ill-conditioned points (mean-weight refinement)
  int ccap = 1;   // candidate-buffer entries per CTA (obs-number-limited search)
  DevBuf<unsigned long long> counters;
  DevBuf<double> st_gues, st_anal, st_gues2, st_anal2, st_infl, st_rtps, st_logp;
  DevBuf<int> st_nobsl;
  DevBuf<double> cb[10];
  DevBuf<int> cb_i;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // host-buffer pipeline of das_letkf: level chunks flow H2D -> analysis -> D2H on three streams
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  std::vector<cudaEvent_t> ev_in, ev_k0, ev_k1;
  std::vector<std::pair<void *, size_t>> pinned;   // caller buffers page-locked by ensure_pinned
  // stats of the last das call
  long long st_points = 0, st_solved = 0, st_fail = 0, st_nobs = 0, st_sweeps = 0, st_refined = 0;
  long long st_phase[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float last_ms = 0.f;
  int last_launches = 0;
  TiledBufs *tiled = nullptr;
  std::map<std::string, void *> ipc_open;   // peer allocations mapped by letkf_b200_peer_open (by handle bytes)
  bool attr_gemm_big = false, attr_gemm_small = false;   // dynamic-shared-memory opt-in done on this handle's device
  // pre-search pipeline (presearch_kernel): three pools (chunk index mod 3), the search runs two chunks ahead of the solver
  cudaStream_t s_search = nullptr;
  std::vector<cudaEvent_t> ev_s;
  static constexpr int kPools = 3;   // pooled local lists of three level chunks: the search runs two chunks ahead of the solver
  DevBuf<int> pl_n[kPools], pl_iob[kPools], ps_l_iob;
  DevBuf<long long> pl_off[kPools];
  DevBuf<double> pl_rdiag[kPools], pl_rloc[kPools], ps_l_rdiag, ps_l_rloc, ps_l_cnd;
  DevBuf<unsigned> ps_l_cpk;
  DevBuf<unsigned long long> ps_counters;   // [0..15] work counter block, [18] redo count, [20..22] pool cursors
  int ps_grid = 0;
  DevBuf<long long> redo_list;
  // scratch of set_obs
  DevBuf<int> so_ic, so_key, so_count, so_fill, so_tmp, so_use, so_keep, so_pos, so_ic0, so_kept;
  DevBuf<double> so_vc0;
  DevBuf<double> so_ri, so_rj, so_vc, so_err, so_val, so_ens;
};

#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                       \
      return LETKF_B200_ECUDA;                                                           \
    }                                                                                    \
  } while (0)

namespace {

int fail(letkf_b200_handle *h, int code, const char *msg) {
  h->err = msg;
  return code;
}

// letkf_tools.f90:130-163 -- group variables with identical var_local rows
void setup_var_groups(letkf_b200_handle *h) {
  const letkf_b200_config &c = h->cfg;
  const int nv = c.nv3d + c.nv2d;
  int n2nc[kMaxNV], n2n[kMaxNV], n2nc_max = 1;
  n2nc[0] = 1;
  n2n[0] = 1;
  for (int n = 2; n <= nv; ++n) {
    bool found = false;
    for (int i = 1; i <= n2nc_max; ++i) {
      const int ref = n2nc[i - 1];   // as written in the reference (:146)
      double md = 0.0;
      for (int iv = 0; iv < LETKF_B200_NID_VARLOCAL; ++iv)
        md = std::max(md, std::fabs(c.VAR_LOCAL[iv][ref - 1] - c.VAR_LOCAL[iv][n - 1]));
      if (md < std::numeric_limits<double>::min()) {
        n2nc[n - 1] = n2nc[i - 1];
        n2n[n - 1] = n2n[n2nc[n - 1] - 1];
        found = true;
        break;
      }
    }
    if (!found) {
      ++n2nc_max;
      n2nc[n - 1] = n2nc_max;
      n2n[n - 1] = n;
    }
  }
  h->nvgroup = n2nc_max;
  for (int n = 0; n < nv; ++n) {
    h->vgroup[n] = n2nc[n] - 1;
    h->vfirst[n] = n2n[n] - 1;
  }
}

// One analysis launch configuration: kernel, block size, dynamic shared memory, resident grid.
struct DasLaunch {
  void (*kern)(DasParams) = nullptr;
  int nt = 0;
  size_t smem = 0;
  long long grid = 0;
  int occ = 1;
};

int plan_common(letkf_b200_handle *h, DasParams &P, DasLaunch &L, int occ, const char *what) {
  if (occ < 1) return fail(h, LETKF_B200_EINVAL, what);
  L.grid = (long long)occ * h->num_sms;   // persistent CTAs: one resident set per SM
  CK(h->l_iob.ensure((size_t)L.grid * P.lcap));
  CK(h->l_rdiag.ensure((size_t)L.grid * P.lcap));
  CK(h->l_rloc.ensure((size_t)L.grid * P.lcap));
  CK(h->l_cnd.ensure((size_t)L.grid * h->ccap));
  CK(h->l_cpk.ensure((size_t)L.grid * h->ccap));
  P.l_iob = h->l_iob.p;
  P.l_rdiag = h->l_rdiag.p;
  P.l_rloc = h->l_rloc.p;
  P.l_cnd = h->l_cnd.p;
  P.l_cpk = h->l_cpk.p;
  P.ccap = h->ccap;
  return LETKF_B200_OK;
}

template <int KC>
int plan_das(letkf_b200_handle *h, DasParams &P, DasLaunch &L) {
  using SC = SizeClass<KC>;
  L.kern = das_kernel<KC>;
  L.nt = SC::NT;
  L.smem = das_smem_bytes(P.k, SC::NT);
  CK(cudaFuncSetAttribute(das_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, das_kernel<KC>, SC::NT, L.smem));
  return plan_common(h, P, L, occ, "das_kernel does not fit on an SM");
}

template <int NB, bool PRE>
int plan_das_ns(letkf_b200_handle *h, DasParams &P, DasLaunch &L) {
  using C = NsCfg<NB>;
  if (P.ldens != C::KP) return fail(h, LETKF_B200_ESTATE, "observation rows were laid out for another ensemble size (call set_obs again)");
  L.kern = das_ns_kernel<NB, PRE>;
  L.nt = C::NT;
  L.smem = das_ns_smem_bytes<NB, PRE>();
  CK(cudaFuncSetAttribute(das_ns_kernel<NB, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, das_ns_kernel<NB, PRE>, C::NT, L.smem));
  if (const char *mo = std::getenv("LETKF_B200_MAXOCC")) occ = std::max(1, std::min(occ, std::atoi(mo)));   // experiments
  L.occ = occ;
  CK(h->m0_scratch.ensure((size_t)std::max(occ, 1) * h->num_sms * C::PSZ));
  P.m0_scratch = h->m0_scratch.p;
  return plan_common(h, P, L, occ, "das_ns_kernel does not fit on an SM");
}
template <bool PRE>
int plan_das_ns_class(letkf_b200_handle *h, int nsc, DasParams &P, DasLaunch &L) {
  if (nsc == 3) return plan_das_ns<3, PRE>(h, P, L);
  if (nsc == 5) return plan_das_ns<5, PRE>(h, P, L);
  if (nsc == 7) return plan_das_ns<7, PRE>(h, P, L);
  if (nsc == 9) return plan_das_ns<9, PRE>(h, P, L);
  return plan_das_ns<13, PRE>(h, P, L);
}

// analyse points [begin, end) on the handle's compute stream, bracketed by the two events
int launch_range(letkf_b200_handle *h, const DasLaunch &L, DasParams P, long long begin, long long end,
                 cudaEvent_t e0, cudaEvent_t e1) {
  P.point_begin = begin;
  P.point_end = end;
  const long long grid = std::min<long long>(L.grid, std::max<long long>(end - begin, 1));
  CK(cudaMemsetAsync(h->counters.p, 0, sizeof(unsigned long long), h->stream));   // work counter
  CK(cudaEventRecord(e0, h->stream));
  L.kern<<<(unsigned)grid, L.nt, L.smem, h->stream>>>(P);
  CK(cudaGetLastError());
  CK(cudaEventRecord(e1, h->stream));
  return LETKF_B200_OK;
}

// Page-lock a caller-owned host buffer once (cached per handle) so that the chunked copies of
// das_letkf overlap with the analysis kernels.  LETKF_B200_PIN=0 leaves pageable memory alone.
void ensure_pinned(letkf_b200_handle *h, void *p, size_t bytes) {
  if (!p || bytes == 0) return;
  const char *pin = std::getenv("LETKF_B200_PIN");
  if (pin && pin[0] == '0') return;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return; }
  if (at.type != cudaMemoryTypeUnregistered) return;
  for (auto it = h->pinned.begin(); it != h->pinned.end();) {   // stale registration of a moved buffer
    char *q = (char *)it->first;
    if ((char *)p < q + it->second && q < (char *)p + bytes) {
      cudaHostUnregister(it->first);
      it = h->pinned.erase(it);
    } else ++it;
  }
  if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess) h->pinned.emplace_back(p, bytes);
  else cudaGetLastError();
}

}  // namespace
#include "tiled_host.cuh"
namespace {

template <int KC>
int launch_core(letkf_b200_handle *h, CoreParams &P) {
  using SC = SizeClass<KC>;
  const size_t smem = core_smem_bytes(P.ne);
  CK(cudaFuncSetAttribute(core_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, core_kernel<KC>, SC::NT, smem));
  if (occ < 1) return fail(h, LETKF_B200_EINVAL, "core_kernel does not fit on an SM");
  const long long grid = std::min<long long>((long long)occ * h->num_sms, std::max(P.npts, 1));
  core_kernel<KC><<<(unsigned)grid, SC::NT, smem, h->stream>>>(P);
  CK(cudaGetLastError());
  return LETKF_B200_OK;
}

}  // namespace

extern "C" {

const char *letkf_b200_build_info(void) { return "libletkf_b200 sm_100a fp64 (CUDA " __DATE__ ")"; }

void letkf_b200_abi_sizes(int32_t sizes[4]) {
  sizes[0] = (int32_t)sizeof(letkf_b200_config);
  sizes[1] = (int32_t)sizeof(letkf_b200_ctype_info);
  sizes[2] = (int32_t)sizeof(letkf_b200_obs);
  sizes[3] = (int32_t)sizeof(letkf_b200_das_args);
}

void letkf_b200_config_defaults(letkf_b200_config *c) {
  // scale/common/common_nml.f90 defaults (:40-46, :109-142, :160-229, :264)
  static const double min_spacing[LETKF_B200_NOBTYPE] = {
      300.0e3, 100.0e3, 100.0e3, 150.0e3, 300.0e3, 150.0e3, 150.0e3, 100.0e3, 150.0e3, 150.0e3, 150.0e3, 150.0e3,
      150.0e3, 150.0e3, 150.0e3, 150.0e3, 300.0e3, 150.0e3, 150.0e3, 150.0e3, 150.0e3, 1.0e3,   15.0e3,  1000.0e3};
  std::memset(c, 0, sizeof(*c));
  c->MEMBER = 3;
  c->nv3d = 11;
  c->IHALO = c->JHALO = 2;
  c->DX = c->DY = 1.0;
  c->iv3d_p = 5;
  c->iv3d_q = 6;
  c->iv3d_qg = 11;
  c->INFL_MUL = 1.0;
  c->INFL_MUL_MIN = -1.0;
  c->Q_SPRD_MAX = -1.0;
  for (int t = 0; t < LETKF_B200_NOBTYPE; ++t) {
    c->HORI_LOCAL[t] = -1.0;
    c->VERT_LOCAL[t] = -1.0;
    c->MAX_NOBS_PER_GRID[t] = -1;
    c->OBS_MIN_SPACING[t] = min_spacing[t];
    c->OBS_SORT_GRID_SPACING[t] = -1.0;
  }
  c->HORI_LOCAL[0] = 500.0e3;
  c->VERT_LOCAL[0] = 0.4;
  c->VERT_LOCAL[21] = 1000.0;
  c->MAX_NOBS_PER_GRID[0] = 0;
  c->OBS_SORT_GRID_SPACING[0] = 0.0;
  c->HORI_LOCAL_RADAR_OBSNOREF = c->HORI_LOCAL_RADAR_VR = c->VERT_LOCAL_RADAR_VR = -1.0;
  c->VERT_LOCAL_RAIN_BASE = 85000.0;
  c->MAX_NOBS_PER_GRID_CRITERION = 1;
  for (int iv = 0; iv < LETKF_B200_NID_VARLOCAL; ++iv)
    for (int n = 0; n < LETKF_B200_MAX_NV; ++n) c->VAR_LOCAL[iv][n] = 1.0;
  c->RADAR_ZMAX = 99.0e3;
  // letkf_obs.f90:27-28: default-REAL literals widened to double
  c->dist_zero_fac = (double)3.651483717f;
  c->dist_zero_fac_square = (double)13.33333333f;
}

void letkf_b200_config_resolve(letkf_b200_config *c) {   // common_nml.f90:741-775
  for (int t = 1; t < LETKF_B200_NOBTYPE; ++t) {
    if (c->HORI_LOCAL[t] < 0.0) c->HORI_LOCAL[t] = c->HORI_LOCAL[0];
    if (c->VERT_LOCAL[t] < 0.0) c->VERT_LOCAL[t] = c->VERT_LOCAL[0];
    if (c->MAX_NOBS_PER_GRID[t] < 0) c->MAX_NOBS_PER_GRID[t] = c->MAX_NOBS_PER_GRID[0];
    if (c->OBS_MIN_SPACING[t] <= 0.0) c->OBS_MIN_SPACING[t] = c->OBS_MIN_SPACING[0];
    if (c->OBS_SORT_GRID_SPACING[t] < 0.0) c->OBS_SORT_GRID_SPACING[t] = c->OBS_SORT_GRID_SPACING[0];
  }
  if (c->HORI_LOCAL_RADAR_OBSNOREF < 0.0) c->HORI_LOCAL_RADAR_OBSNOREF = c->HORI_LOCAL[21];
  if (c->HORI_LOCAL_RADAR_VR < 0.0) c->HORI_LOCAL_RADAR_VR = c->HORI_LOCAL[21];
  if (c->VERT_LOCAL_RADAR_VR < 0.0) c->VERT_LOCAL_RADAR_VR = c->VERT_LOCAL[21];
}

int letkf_b200_create(const letkf_b200_config *cfg, int device, letkf_b200_handle **out) {
  if (!cfg || !out) return LETKF_B200_EINVAL;
  *out = nullptr;
  letkf_b200_handle *h = new letkf_b200_handle();
  h->cfg = *cfg;
  h->device = device;
  const letkf_b200_config &c = h->cfg;
  auto bad = [&](const char *m) {
    std::fprintf(stderr, "letkf_b200_create: %s\n", m);
    delete h;
    return LETKF_B200_EINVAL;
  };
  if (c.MEMBER < 2 || c.MEMBER > LETKF_B200_MAX_MEMBER) return bad("MEMBER must be in [2, LETKF_B200_MAX_MEMBER]");
  if (c.nv3d < 1 || c.nv3d + c.nv2d > kMaxNV - 2) return bad("nv3d + nv2d must be in [1, 14]");
  if (c.nlon < 1 || c.nlat < 1 || c.nlev < 1) return bad("nlon/nlat/nlev must be positive");
  if (c.MAX_NOBS_PER_GRID_CRITERION < 1 || c.MAX_NOBS_PER_GRID_CRITERION > 3) return bad("Unsupported MAX_NOBS_PER_GRID_CRITERION");
  if (c.iv3d_p < 1 || c.iv3d_p > c.nv3d) return bad("iv3d_p out of range");
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    std::fprintf(stderr, "letkf_b200_create: cudaSetDevice(%d): %s (no CPU fallback exists)\n", device,
                 cudaGetErrorString(e));
    delete h;
    return LETKF_B200_ECUDA;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    delete h;
    return LETKF_B200_ECUDA;
  }
  h->num_sms = prop.multiProcessorCount;
  cudaEventCreate(&h->ev0);
  cudaEventCreate(&h->ev1);
  cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking);
  {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    cudaStreamCreateWithPriority(&h->s_search, cudaStreamNonBlocking, hi);
  }
  if (h->counters.ensure(16) != cudaSuccess || h->ps_counters.ensure(24) != cudaSuccess) {
    delete h;
    return LETKF_B200_ECUDA;
  }
  setup_var_groups(h);
  *out = h;
  return LETKF_B200_OK;
}

int letkf_b200_destroy(letkf_b200_handle *h) {
  if (!h) return LETKF_B200_OK;
  cudaSetDevice(h->device);
  h->rig1.release(); h->rjg1.release(); h->hgt1.release();
  h->d_tables.release(); h->rec.release(); h->bstart.release(); h->s2o.release();
  h->sval.release(); h->sens.release(); h->vlfac_groups.release(); h->vlfac_one.release();
  h->l_iob.release(); h->l_rdiag.release(); h->l_rloc.release(); h->l_cnd.release(); h->l_cpk.release(); h->counters.release(); h->m0_scratch.release();
  h->st_gues.release(); h->st_anal.release(); h->st_gues2.release(); h->st_anal2.release();
  h->st_infl.release(); h->st_rtps.release(); h->st_logp.release(); h->st_nobsl.release();
  for (auto &b : h->cb) b.release();
  h->cb_i.release();
  if (h->tiled) { h->tiled->release(); delete h->tiled; h->tiled = nullptr; }
  for (int b = 0; b < letkf_b200_handle::kPools; ++b) {
    h->pl_n[b].release(); h->pl_iob[b].release(); h->pl_off[b].release(); h->pl_rdiag[b].release(); h->pl_rloc[b].release();
  }
  h->ps_l_iob.release(); h->ps_l_rdiag.release(); h->ps_l_rloc.release(); h->ps_l_cnd.release(); h->ps_l_cpk.release();
  h->ps_counters.release(); h->redo_list.release();
  h->so_use.release(); h->so_keep.release(); h->so_pos.release(); h->so_ic0.release(); h->so_kept.release(); h->so_vc0.release();
  h->so_ic.release(); h->so_key.release(); h->so_count.release(); h->so_fill.release(); h->so_tmp.release();
  h->so_ri.release(); h->so_rj.release(); h->so_vc.release(); h->so_err.release(); h->so_val.release(); h->so_ens.release();
  for (cudaEvent_t e : h->ev_s) cudaEventDestroy(e);
  if (h->s_search) cudaStreamDestroy(h->s_search);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (auto *v : {&h->ev_in, &h->ev_k0, &h->ev_k1})
    for (cudaEvent_t e : *v) cudaEventDestroy(e);
  if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
  if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
  for (auto &pr : h->pinned) cudaHostUnregister(pr.first);
  for (auto &kv : h->ipc_open) cudaIpcCloseMemHandle(kv.second);
  cudaGetLastError();
  delete h;
  return LETKF_B200_OK;
}

const char *letkf_b200_last_error(const letkf_b200_handle *h) { return h ? h->err.c_str() : "null handle"; }

int letkf_b200_set_stream(letkf_b200_handle *h, void *s) {
  if (!h) return LETKF_B200_EINVAL;
  h->stream = (cudaStream_t)s;
  return LETKF_B200_OK;
}

int letkf_b200_set_grid(letkf_b200_handle *h, int nij1, const double *rig1, const double *rjg1,
                        const double *hgt1, int mem_space) {
  if (!h || nij1 < 1 || !rig1 || !rjg1 || !hgt1) return LETKF_B200_EINVAL;
  CK(cudaSetDevice(h->device));
  const cudaMemcpyKind kind = mem_space == LETKF_B200_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  CK(h->rig1.ensure(nij1));
  CK(h->rjg1.ensure(nij1));
  CK(h->hgt1.ensure((size_t)nij1 * h->cfg.nlev));
  CK(cudaMemcpyAsync(h->rig1.p, rig1, sizeof(double) * nij1, kind, h->stream));
  CK(cudaMemcpyAsync(h->rjg1.p, rjg1, sizeof(double) * nij1, kind, h->stream));
  CK(cudaMemcpyAsync(h->hgt1.p, hgt1, sizeof(double) * (size_t)nij1 * h->cfg.nlev, kind, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->nij1 = nij1;
  return LETKF_B200_OK;
}

}  // extern "C"

// set_letkf_obs twin.  device = false: every array of `obs` is a host array of accepted observations (the reference's
// obsda after its QC filter).  device = true: device arrays straight out of the observation operator / departure-QC
// kernels, with their QC flags; the filter qc == iqc_good (letkf_obs.f90:752, 791), the combined-type lookup and the
// vertical coordinate are done on the device, nothing but a 1.5 KB occupancy table crosses PCIe.
static int set_obs_impl(letkf_b200_handle *h, const letkf_b200_obs *obs, const int32_t *qc, bool device, int32_t *nkept_out) {
  if (!h || !obs || obs->nobs < 0) return LETKF_B200_EINVAL;
  CK(cudaSetDevice(h->device));
  const letkf_b200_config &c = h->cfg;
  int nobs = obs->nobs;
  const int nobs_in = obs->nobs;
  const int need = c.DET_RUN ? c.MEMBER + 1 : c.MEMBER;
  if (nobs > 0 && obs->nensobs < need) return fail(h, LETKF_B200_EINVAL, "nensobs < MEMBER (+1 with DET_RUN)");
  if (nobs >= (1 << 28)) return fail(h, LETKF_B200_EINVAL, "more than 2^28 observations");
  h->obs_set = false;
  // ---- ctype table (letkf_obs.f90:300-342) --------------------------------------------------
  bool use[LETKF_B200_NID_OBS][LETKF_B200_NOBTYPE];
  std::memset(use, 0, sizeof(use));
  DevBuf<int> &d_use = h->so_use;   // [NID_OBS * NOBTYPE] occupancy, [.. + 1] error flags, then the lookup tables
  if (!device) {
    for (int n = 0; n < nobs; ++n) {
      const int u = uid_obs(obs->elm[n]);
      if (u < 1 || obs->typ[n] < 1 || obs->typ[n] > LETKF_B200_NOBTYPE) return fail(h, LETKF_B200_EINVAL, "unknown obs elm/typ");
      use[u - 1][obs->typ[n] - 1] = true;
    }
  } else {
    const int nu = LETKF_B200_NID_OBS * LETKF_B200_NOBTYPE;
    CK(d_use.ensure(2 * nu + kMaxCtype + 8));
    CK(cudaMemsetAsync(d_use.p, 0, sizeof(int) * (nu + 1), h->stream));
    if (nobs > 0)
      obs_use_kernel<<<(nobs + 255) / 256, 256, 0, h->stream>>>(nobs, LETKF_B200_NOBTYPE, obs->elm, obs->typ, qc, d_use.p, d_use.p + nu);
    std::vector<int> hu(nu + 1);
    CK(cudaMemcpyAsync(hu.data(), d_use.p, sizeof(int) * (nu + 1), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (hu[nu] & 1) return fail(h, LETKF_B200_EINVAL, "unknown obs elm/typ");
    for (int u = 0; u < LETKF_B200_NID_OBS; ++u)
      for (int t = 0; t < LETKF_B200_NOBTYPE; ++t) use[u][t] = hu[u * LETKF_B200_NOBTYPE + t] != 0;
  }
  int ctype_elmtyp[LETKF_B200_NID_OBS][LETKF_B200_NOBTYPE];
  std::memset(ctype_elmtyp, 0, sizeof(ctype_elmtyp));
  SearchTables &T = h->tables;
  std::memset(&T, 0, sizeof(T));
  h->ctinfo.clear();
  int nct = 0, boff = 0;
  for (int ityp = 1; ityp <= LETKF_B200_NOBTYPE; ++ityp)
    for (int ielm_u = 1; ielm_u <= LETKF_B200_NID_OBS; ++ielm_u) {
      if (!use[ielm_u - 1][ityp - 1]) continue;
      if (nct >= kMaxCtype) return fail(h, LETKF_B200_EINVAL, "too many combined obs types");
      ctype_elmtyp[ielm_u - 1][ityp - 1] = nct + 1;
      const int elm = ELEM_UID[ielm_u - 1];
      CtypeDev &d = T.ct[nct];
      letkf_b200_ctype_info info;
      std::memset(&info, 0, sizeof(info));
      d.hori_loc = (elm == ID_RE0) ? c.HORI_LOCAL_RADAR_OBSNOREF
                   : (elm == ID_VR) ? c.HORI_LOCAL_RADAR_VR : c.HORI_LOCAL[ityp - 1];
      d.vert_loc = (elm == ID_VR) ? c.VERT_LOCAL_RADAR_VR : c.VERT_LOCAL[ityp - 1];
      // sorting mesh (letkf_obs.f90:660-695)
      double target;
      if (c.OBS_SORT_GRID_SPACING[ityp - 1] > 0) target = c.OBS_SORT_GRID_SPACING[ityp - 1];
      else if (c.MAX_NOBS_PER_GRID[ityp - 1] > 0)
        target = 0.1 * std::sqrt((double)c.MAX_NOBS_PER_GRID[ityp - 1]) * c.OBS_MIN_SPACING[ityp - 1];
      else target = d.hori_loc * c.dist_zero_fac / 6.0;
      d.ngrd_i = std::min((int)std::ceil(c.DX * (double)c.nlon / target), c.nlon);
      d.ngrd_j = std::min((int)std::ceil(c.DY * (double)c.nlat / target), c.nlat);
      d.grdspc_i = c.DX * (double)c.nlon / (double)d.ngrd_i;
      d.grdspc_j = c.DY * (double)c.nlat / (double)d.ngrd_j;
      d.ngrdsch_i = (int)std::ceil(d.hori_loc * c.dist_zero_fac / d.grdspc_i);
      d.ngrdsch_j = (int)std::ceil(d.hori_loc * c.dist_zero_fac / d.grdspc_j);
      d.ngrdext_i = d.ngrd_i + d.ngrdsch_i * 2;
      d.ngrdext_j = d.ngrd_j + d.ngrdsch_j * 2;
      d.boff = boff;
      boff += d.ngrdext_i * d.ngrdext_j;
      d.elm_u = ielm_u;
      d.typ = ityp;
      d.varlocal = uid_obs_varlocal(elm) - 1;
      d.vconst = 0.0;
      // vertical coordinate mode, same precedence as obs_local_cal (letkf_tools.f90:1852-1866)
      if (d.vert_loc == 0.0) d.vmode = 0;
      else if (elm == ID_PS) d.vmode = 1;
      else if (elm == ID_RAIN) { d.vmode = 2; d.vconst = std::log(c.VERT_LOCAL_RAIN_BASE); }
      else if (ityp == 22) d.vmode = 3;
      else d.vmode = 1;
      {
        const double guard = 1.0 + 9.094947017729282e-13;   // 1 + 2^-40
        d.vmax = c.dist_zero_fac * std::fabs(d.vert_loc) * guard;
        d.hmax2 = (c.dist_zero_fac * d.hori_loc) * (c.dist_zero_fac * d.hori_loc) * guard;
        d.ih2 = 1.0 / (d.hori_loc * d.hori_loc);
        d.iv2 = (d.vmode == 0) ? 0.0 : 1.0 / (d.vert_loc * d.vert_loc);
      }
      info.elm = elm; info.elm_u = ielm_u; info.typ = ityp;
      info.ngrd_i = d.ngrd_i; info.ngrd_j = d.ngrd_j; info.ngrdsch_i = d.ngrdsch_i; info.ngrdsch_j = d.ngrdsch_j;
      info.ngrdext_i = d.ngrdext_i; info.ngrdext_j = d.ngrdext_j;
      info.hori_loc = d.hori_loc; info.vert_loc = d.vert_loc; info.grdspc_i = d.grdspc_i; info.grdspc_j = d.grdspc_j;
      h->ctinfo.push_back(info);
      ++nct;
    }
  h->nctype = nct;
  h->nbuckets = boff;
  T.nctype = nct;
  T.criterion = c.MAX_NOBS_PER_GRID_CRITERION;
  T.IHALO = c.IHALO; T.JHALO = c.JHALO; T.nlon = c.nlon; T.nlat = c.nlat;
  T.DX = c.DX; T.DY = c.DY; T.dzf = c.dist_zero_fac; T.dzf2 = c.dist_zero_fac_square;
  // ---- per-obs ctype and vertical coordinate (host libm: same log() as a CPU run) ---------------
  std::vector<int> ic_of(device ? 0 : nobs);
  std::vector<double> vc(device ? 0 : nobs);
  for (int n = 0; n < (device ? 0 : nobs); ++n) {
    const int ic = ctype_elmtyp[uid_obs(obs->elm[n]) - 1][obs->typ[n] - 1] - 1;
    ic_of[n] = ic;
    const CtypeDev &d = T.ct[ic];
    if (d.vmode == 1) {   // localised in ln p: a non-positive pressure would put NaN into every distance test
      const double pr = (obs->elm[n] == ID_PS) ? obs->dat[n] : obs->lev[n];
      if (!(pr > 0.0)) return fail(h, LETKF_B200_EINVAL, "non-positive pressure in an observation localised in ln p");
    }
    if (d.vmode == 3) vc[n] = obs->lev[n];
    else if (obs->elm[n] == ID_PS) vc[n] = std::log(obs->dat[n]);
    else vc[n] = std::log(obs->lev[n]);
  }
  // ---- device bucket sort ------------------------------------------------------------------------
  h->nensobs = obs->nensobs;
  h->ldens = ns_row_doubles(c.MEMBER);   // [ensval(1..k) | dep | depd | 0..]
  h->nobstotal = nobs;
  // scratch of the sort: kept in the handle (grow-only) -- eleven cudaMalloc/cudaFree pairs per call cost
  // 100-500 ms inside a process that holds ~100 GB of device memory (cudaFree synchronises the device)
  DevBuf<int> &d_ic = h->so_ic, &d_key = h->so_key, &d_count = h->so_count, &d_fill = h->so_fill, &d_tmp = h->so_tmp;
  DevBuf<double> &d_ri = h->so_ri, &d_rj = h->so_rj, &d_vc = h->so_vc, &d_err = h->so_err, &d_val = h->so_val, &d_ens = h->so_ens;
  CK(h->d_tables.ensure(1));
  CK(cudaMemcpyAsync(h->d_tables.p, &T, sizeof(T), cudaMemcpyHostToDevice, h->stream));
  CK(h->bstart.ensure((size_t)boff + 1));
  CK(h->s2o.ensure(nobs));
  CK(h->rec.ensure(nobs));
  CK(h->sval.ensure(nobs));
  CK(h->sens.ensure((size_t)nobs * h->ldens));
  CK(d_count.ensure(boff + 1));
  CK(d_fill.ensure(boff + 1));
  CK(cudaMemsetAsync(d_count.p, 0, sizeof(int) * (boff + 1), h->stream));
  CK(cudaMemsetAsync(d_fill.p, 0, sizeof(int) * (boff + 1), h->stream));
  if (device && nobs > 0) {
    // combined type, vertical coordinate and the qc == 0 compaction on the device
    const int nu = LETKF_B200_NID_OBS * LETKF_B200_NOBTYPE;
    std::vector<int> tab(nu + kMaxCtype, 0);
    for (int u = 0; u < LETKF_B200_NID_OBS; ++u)
      for (int t = 0; t < LETKF_B200_NOBTYPE; ++t) tab[u * LETKF_B200_NOBTYPE + t] = ctype_elmtyp[u][t];
    for (int ic = 0; ic < nct; ++ic) tab[nu + ic] = T.ct[ic].vmode;
    int *d_tab = d_use.p + nu + 1, *d_err = d_use.p + nu;
    CK(cudaMemcpyAsync(d_tab, tab.data(), sizeof(int) * tab.size(), cudaMemcpyHostToDevice, h->stream));
    CK(h->so_keep.ensure(nobs)); CK(h->so_pos.ensure((size_t)nobs + 1)); CK(h->so_ic0.ensure(nobs)); CK(h->so_vc0.ensure(nobs));
    obs_prepare_kernel<<<(nobs + 255) / 256, 256, 0, h->stream>>>(nobs, LETKF_B200_NOBTYPE, obs->elm, obs->typ, qc, obs->lev, obs->dat,
                                                                d_tab, d_tab + nu, h->so_keep.p, h->so_ic0.p, h->so_vc0.p, d_err);
    exclusive_scan_kernel<<<1, 256, 0, h->stream>>>(h->so_keep.p, h->so_pos.p, nobs);
    int hk[2] = {0, 0};
    CK(cudaMemcpyAsync(&hk[0], h->so_pos.p + nobs, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(&hk[1], d_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (hk[1] & 2) return fail(h, LETKF_B200_EINVAL, "non-positive pressure in an observation localised in ln p");
    nobs = hk[0];
    h->nobstotal = nobs;
    CK(h->s2o.ensure(std::max(nobs, 1))); CK(h->rec.ensure(std::max(nobs, 1))); CK(h->sval.ensure(std::max(nobs, 1)));
    CK(h->sens.ensure((size_t)std::max(nobs, 1) * h->ldens));
    CK(h->so_kept.ensure(std::max(nobs, 1)));
  }
  if (nobs_in > 0 && nobs == 0 && device) {   // everything rejected: empty tables
    exclusive_scan_kernel<<<1, 256, 0, h->stream>>>(d_count.p, h->bstart.p, boff);
    CK(cudaGetLastError());
  } else
  if (nobs > 0) {
    CK(d_ic.ensure(nobs)); CK(d_key.ensure(nobs)); CK(d_tmp.ensure(nobs));
    CK(d_ri.ensure(nobs)); CK(d_rj.ensure(nobs)); CK(d_vc.ensure(nobs)); CK(d_err.ensure(nobs));
    CK(d_val.ensure(nobs)); CK(d_ens.ensure((size_t)nobs * obs->nensobs));
    const size_t nb = sizeof(double) * nobs;
    if (device) {
      obs_compact_kernel<<<nobs_in, 64, 0, h->stream>>>(nobs_in, obs->nensobs, h->so_keep.p, h->so_pos.p, h->so_ic0.p, h->so_vc0.p,
                                                       obs->ri, obs->rj, obs->err, obs->val, obs->ensval, d_ic.p, d_vc.p, d_ri.p,
                                                       d_rj.p, d_err.p, d_val.p, d_ens.p, h->so_kept.p);
      CK(cudaGetLastError());
    } else {
    CK(cudaMemcpyAsync(d_ic.p, ic_of.data(), sizeof(int) * nobs, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_ri.p, obs->ri, nb, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_rj.p, obs->rj, nb, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_vc.p, vc.data(), nb, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_err.p, obs->err, nb, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_val.p, obs->val, nb, cudaMemcpyHostToDevice, h->stream));
    // the observation-space ensemble is the one large table (C3: 1.2 GB): page-lock the caller's buffer once
    // (cached per handle) so that the copy runs at PCIe speed instead of the pageable-memory staging rate
    if (nb * obs->nensobs >= ((size_t)32 << 20)) ensure_pinned(h, const_cast<double *>(obs->ensval), nb * obs->nensobs);
    CK(cudaMemcpyAsync(d_ens.p, obs->ensval, nb * obs->nensobs, cudaMemcpyHostToDevice, h->stream));
    }
    const int tb = 256, gb = (nobs + tb - 1) / tb;
    bucket_key_kernel<<<gb, tb, 0, h->stream>>>(h->d_tables.p, nobs, d_ic.p, d_ri.p, d_rj.p, d_key.p, d_count.p);
    exclusive_scan_kernel<<<1, 256, 0, h->stream>>>(d_count.p, h->bstart.p, boff);
    bucket_scatter_kernel<<<gb, tb, 0, h->stream>>>(nobs, d_key.p, h->bstart.p, d_fill.p, d_tmp.p);
    bucket_rank_kernel<<<gb, tb, 0, h->stream>>>(nobs, d_key.p, h->bstart.p, d_tmp.p, h->s2o.p);
    obs_gather_kernel<<<nobs, 64, 0, h->stream>>>(nobs, obs->nensobs, c.MEMBER, h->ldens, h->s2o.p, d_ri.p, d_rj.p, d_vc.p,
                                                  d_err.p, d_val.p, d_ens.p, h->rec.p, h->sval.p, h->sens.p);
    CK(cudaGetLastError());
  } else {
    exclusive_scan_kernel<<<1, 256, 0, h->stream>>>(d_count.p, h->bstart.p, boff);
    CK(cudaGetLastError());
  }
  h->h_bstart.resize((size_t)boff + 1);
  h->h_s2o.resize(nobs);
  CK(cudaMemcpyAsync(h->h_bstart.data(), h->bstart.p, sizeof(int) * ((size_t)boff + 1), cudaMemcpyDeviceToHost, h->stream));
  if (nobs > 0) CK(cudaMemcpyAsync(h->h_s2o.data(), h->s2o.p, sizeof(int) * nobs, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int ic = 0; ic < nct; ++ic) {
    const CtypeDev &d = T.ct[ic];
    const int b0 = h->h_bstart[d.boff], b1 = h->h_bstart[d.boff + d.ngrdext_i * d.ngrdext_j];
    T.ct[ic].tot = b1 - b0;
    h->ctinfo[ic].tot_ext = b1 - b0;
    h->ctinfo[ic].ac_begin = b0;
  }
  // ---- merged obs-number budgets (letkf_tools.f90:167-192) -------------------------------------
  std::vector<int> n_merge(nct, 1);
  T.ngroup = 0;
  int maxl = 0;
  auto merge_id = [&](int ic) { return (T.ct[ic].typ == 22 && (T.ct[ic].elm_u == 9 || T.ct[ic].elm_u == 10)) ? 1 : 0; };
  for (int ic = 0; ic < nct; ++ic) {
    if (n_merge[ic] == 0) continue;
    GroupDev &G = T.grp[T.ngroup++];
    G.n = 1;
    G.ic[0] = ic;
    G.limit = c.MAX_NOBS_PER_GRID[T.ct[ic].typ - 1];
    if (merge_id(ic) > 0)
      for (int ic2 = ic + 1; ic2 < nct; ++ic2)
        if (merge_id(ic2) == merge_id(ic)) {
          if (G.n >= kMaxMerge) return fail(h, LETKF_B200_EINVAL, "too many merged obs types");
          G.ic[G.n++] = ic2;
          n_merge[ic] += 1;
          n_merge[ic2] = 0;
        }
    int tot = 0;
    for (int m = 0; m < G.n; ++m) tot += T.ct[G.ic[m]].tot;
    maxl += (G.limit > 0) ? std::min(G.limit, tot) : tot;
  }
  for (int ic = 0; ic < nct; ++ic) h->ctinfo[ic].n_merge = n_merge[ic];
  h->maxl = (std::max(maxl, 1) + 3) & ~3;   // a multiple of four: the solver pads its lists to four-row steps
  {   // Candidate buffer entries per CTA.  Slices are laid out by candidate position (search.cuh), so the
      // buffer must span the candidates of one search rectangle although only the survivors are
      // written: 32 K entries cover a radar rectangle of ~3 search increments; pages never touched cost
      // nothing.  Without an obs-number limit the scan is windowed and 4 K entries are plenty.
    int nmax = 0;
    for (int g = 0; g < T.ngroup; ++g) nmax = std::max(nmax, T.grp[g].limit);
    h->ccap = nmax > 0 ? 32768 : 4096;
    const char *ce = std::getenv("LETKF_B200_CAND_CAP");   // test hook: force the re-scan fallback
    T.cand_cap_limit = ce ? std::max(0, std::atoi(ce)) : h->ccap;
  }
  h->radar_only = true;   // letkf_tools.f90:197-203
  for (int ic = 0; ic < nct; ++ic)
    if (T.ct[ic].typ != 22) h->radar_only = false;
  CK(cudaMemcpyAsync(h->d_tables.p, &T, sizeof(T), cudaMemcpyHostToDevice, h->stream));
  // ---- variable-localisation factors -------------------------------------------------------------
  const int nv = c.nv3d + c.nv2d;
  h->h_vlfac.assign((size_t)nv * std::max(nct, 1), 1.0);
  for (int n = 0; n < nv; ++n)
    for (int ic = 0; ic < nct; ++ic) h->h_vlfac[(size_t)n * nct + ic] = c.VAR_LOCAL[T.ct[ic].varlocal][n];
  std::vector<double> vg((size_t)h->nvgroup * std::max(nct, 1), 1.0);
  for (int n = nv - 1; n >= 0; --n)   // representative = first variable of each group
    for (int ic = 0; ic < nct; ++ic) vg[(size_t)h->vgroup[n] * nct + ic] = h->h_vlfac[(size_t)n * nct + ic];
  CK(h->vlfac_groups.ensure(vg.size()));
  CK(cudaMemcpyAsync(h->vlfac_groups.p, vg.data(), sizeof(double) * vg.size(), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->obs_set = true;
  if (nkept_out) *nkept_out = nobs;
  return LETKF_B200_OK;
}

extern "C" {

int letkf_b200_set_obs(letkf_b200_handle *h, const letkf_b200_obs *obs) { return set_obs_impl(h, obs, nullptr, false, nullptr); }

int letkf_b200_set_obs_device(letkf_b200_handle *h, const letkf_b200_obs *obs, const int32_t *qc, int32_t *nkept) {
  return set_obs_impl(h, obs, qc, true, nkept);
}

int letkf_b200_get_kept_index(const letkf_b200_handle *h, int32_t *kept) {
  if (!h || !h->obs_set || !kept) return LETKF_B200_ESTATE;
  if (h->so_kept.n < (size_t)std::max(h->nobstotal, 1)) return LETKF_B200_ESTATE;   // not a set_obs_device table
  if (h->nobstotal > 0 && cudaMemcpy(kept, h->so_kept.p, sizeof(int) * h->nobstotal, cudaMemcpyDeviceToHost) != cudaSuccess)
    return LETKF_B200_ECUDA;
  return LETKF_B200_OK;
}

int letkf_b200_abi_size_qc(void) { return (int)sizeof(letkf_b200_qc_config); }

void letkf_b200_qc_config_defaults(letkf_b200_qc_config *q) {   // common_nml.f90:129-137, 248-259
  std::memset(q, 0, sizeof(*q));
  q->GROSS_ERROR = 5.0;
  q->GROSS_ERROR_RAIN = q->GROSS_ERROR_RADAR_REF = q->GROSS_ERROR_RADAR_VR = q->GROSS_ERROR_RADAR_PRH = -1.0;
  q->GROSS_ERROR_TCX = q->GROSS_ERROR_TCY = q->GROSS_ERROR_TCP = -1.0;
  q->RADAR_REF_THRES_DBZ = 15.0;
  q->USE_RADAR_REF = 1;
  q->USE_RADAR_VR = 1;
  q->MIN_RADAR_REF_MEMBER = 1;
  q->MIN_RADAR_REF_MEMBER_OBSREF = 1;
}

int letkf_b200_obs_departure_qc(letkf_b200_handle *h, const letkf_b200_qc_config *q, int nobs, int nensobs,
                                const int32_t *elm, const double *dat, const double *err, int32_t *qc, double *ensval,
                                double *val, int mem_space) {
  if (!h || !q || nobs < 0) return LETKF_B200_EINVAL;
  if (nobs == 0) return LETKF_B200_OK;
  if (!elm || !dat || !err || !qc || !ensval || !val) return LETKF_B200_EINVAL;
  const letkf_b200_config &c = h->cfg;
  if (nensobs < (c.DET_RUN ? c.MEMBER + 1 : c.MEMBER)) return fail(h, LETKF_B200_EINVAL, "nensobs < MEMBER (+1 with DET_RUN)");
  CK(cudaSetDevice(h->device));
  const bool host = mem_space != LETKF_B200_MEM_DEVICE;
  auto ge = [&](double v) { return v < 0.0 ? q->GROSS_ERROR : v; };   // (common_nml.f90:619-642)
  QcParams P;
  std::memset(&P, 0, sizeof(P));
  P.ge = q->GROSS_ERROR; P.ge_rain = ge(q->GROSS_ERROR_RAIN); P.ge_ref = ge(q->GROSS_ERROR_RADAR_REF);
  P.ge_vr = ge(q->GROSS_ERROR_RADAR_VR); P.ge_prh = ge(q->GROSS_ERROR_RADAR_PRH); P.ge_tcx = ge(q->GROSS_ERROR_TCX);
  P.ge_tcy = ge(q->GROSS_ERROR_TCY); P.ge_tcp = ge(q->GROSS_ERROR_TCP); P.ref_thres = q->RADAR_REF_THRES_DBZ;
  P.use_ref = q->USE_RADAR_REF; P.use_vr = q->USE_RADAR_VR; P.min_mem = q->MIN_RADAR_REF_MEMBER;
  P.min_mem_obsref = q->MIN_RADAR_REF_MEMBER_OBSREF;
  P.nobs = nobs; P.nensobs = nensobs; P.member = c.MEMBER; P.det = c.DET_RUN ? 1 : 0;
  const size_t ne = (size_t)nobs * nensobs;
  DevBuf<int> d_elm, d_qc;
  DevBuf<double> d_dat, d_err, d_ens, d_val;
  if (host) {
    CK(d_elm.ensure(nobs)); CK(d_qc.ensure(nobs)); CK(d_dat.ensure(nobs)); CK(d_err.ensure(nobs)); CK(d_ens.ensure(ne)); CK(d_val.ensure(nobs));
    CK(cudaMemcpyAsync(d_elm.p, elm, sizeof(int) * nobs, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_qc.p, qc, sizeof(int) * nobs, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_dat.p, dat, sizeof(double) * nobs, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_err.p, err, sizeof(double) * nobs, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_ens.p, ensval, sizeof(double) * ne, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_val.p, val, sizeof(double) * nobs, cudaMemcpyHostToDevice, h->stream));
    P.elm = d_elm.p; P.qc = d_qc.p; P.dat = d_dat.p; P.err = d_err.p; P.ensval = d_ens.p; P.val = d_val.p;
  } else {
    P.elm = elm; P.qc = qc; P.dat = dat; P.err = err; P.ensval = ensval; P.val = val;
  }
  obs_departure_qc_kernel<<<(unsigned)((nobs + 7) / 8), 256, 0, h->stream>>>(P);
  CK(cudaGetLastError());
  if (host) {
    CK(cudaMemcpyAsync(qc, d_qc.p, sizeof(int) * nobs, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(ensval, d_ens.p, sizeof(double) * ne, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(val, d_val.p, sizeof(double) * nobs, cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  d_elm.release(); d_qc.release(); d_dat.release(); d_err.release(); d_ens.release(); d_val.release();
  return LETKF_B200_OK;
}

int letkf_b200_monit_dep(letkf_b200_handle *h, int nobs, const int32_t *elm, const double *dep, const int32_t *qc,
                         int32_t *nobs_out, double *bias, double *rmse, int mem_space) {
  if (!h || nobs < 0 || !nobs_out || !bias || !rmse) return LETKF_B200_EINVAL;
  if (nobs > 0 && (!elm || !dep || !qc)) return LETKF_B200_EINVAL;
  CK(cudaSetDevice(h->device));
  const bool host = mem_space != LETKF_B200_MEM_DEVICE;
  const int nblocks = std::max(1, std::min((nobs + 255) / 256, h->num_sms * 4));
  CK(h->cb[0].ensure((size_t)nblocks * 48 + 32));
  CK(h->cb_i.ensure(16));
  const int *d_elm = elm, *d_qc = qc;
  const double *d_dep = dep;
  if (host && nobs > 0) {
    CK(h->so_ic.ensure(nobs)); CK(h->so_key.ensure(nobs)); CK(h->so_val.ensure(nobs));
    CK(cudaMemcpyAsync(h->so_ic.p, elm, sizeof(int) * nobs, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->so_key.p, qc, sizeof(int) * nobs, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->so_val.p, dep, sizeof(double) * nobs, cudaMemcpyHostToDevice, h->stream));
    d_elm = h->so_ic.p; d_qc = h->so_key.p; d_dep = h->so_val.p;
  }
  double *part = h->cb[0].p, *d_bias = part + (size_t)nblocks * 48, *d_rmse = d_bias + 16;
  monit_partial_kernel<<<nblocks, 256, 0, h->stream>>>(nobs, d_elm, d_dep, d_qc, part);
  monit_final_kernel<<<1, 32, 0, h->stream>>>(nblocks, part, host ? h->cb_i.p : nobs_out, host ? d_bias : bias,
                                              host ? d_rmse : rmse);
  CK(cudaGetLastError());
  if (host) {
    CK(cudaMemcpyAsync(nobs_out, h->cb_i.p, sizeof(int) * 16, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(bias, d_bias, sizeof(double) * 16, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(rmse, d_rmse, sizeof(double) * 16, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return LETKF_B200_OK;
}

int letkf_b200_obs_info(const letkf_b200_handle *h, int32_t *nobstotal, int32_t *nctype) {
  if (!h || !h->obs_set) return LETKF_B200_ESTATE;
  if (nobstotal) *nobstotal = h->nobstotal;
  if (nctype) *nctype = h->nctype;
  return LETKF_B200_OK;
}
int letkf_b200_get_ctype(const letkf_b200_handle *h, int ic, letkf_b200_ctype_info *out) {
  if (!h || !h->obs_set) return LETKF_B200_ESTATE;
  if (ic < 0 || ic >= h->nctype || !out) return LETKF_B200_EINVAL;
  *out = h->ctinfo[ic];
  return LETKF_B200_OK;
}
int letkf_b200_get_ac_ext(const letkf_b200_handle *h, int ic, int32_t *ac) {
  if (!h || !h->obs_set) return LETKF_B200_ESTATE;
  if (ic < 0 || ic >= h->nctype || !ac) return LETKF_B200_EINVAL;
  const CtypeDev &d = h->tables.ct[ic];
  for (int j = 1; j <= d.ngrdext_j; ++j)
    for (int i = 0; i <= d.ngrdext_i; ++i)
      ac[i + (size_t)(j - 1) * (d.ngrdext_i + 1)] = h->h_bstart[d.boff + (j - 1) * d.ngrdext_i + i];
  return LETKF_B200_OK;
}
int letkf_b200_get_sorted_index(const letkf_b200_handle *h, int32_t *s2o) {
  if (!h || !h->obs_set) return LETKF_B200_ESTATE;
  std::copy(h->h_s2o.begin(), h->h_s2o.end(), s2o);
  return LETKF_B200_OK;
}

int letkf_b200_obs_local(letkf_b200_handle *h, int npts, const double *ri, const double *rj, const double *rlev,
                         const double *rz, int nvar, int32_t *nobsl, int32_t *idx, double *rdiag, double *rloc,
                         int max_out, int mem_space) {
  if (!h || npts < 0 || !nobsl) return LETKF_B200_EINVAL;
  if (!h->obs_set) return fail(h, LETKF_B200_ESTATE, "set_obs has not been called");
  const int nv = h->cfg.nv3d + h->cfg.nv2d;
  if (nvar < 1 || nvar > nv) return fail(h, LETKF_B200_EINVAL, "nvar must be in [1, nv3d+nv2d]");
  CK(cudaSetDevice(h->device));
  const bool host = mem_space != LETKF_B200_MEM_DEVICE;
  const int nct = std::max(h->nctype, 1);
  CK(h->vlfac_one.ensure(nct));
  CK(cudaMemcpyAsync(h->vlfac_one.p, h->h_vlfac.data() + (size_t)(nvar - 1) * h->nctype, sizeof(double) * h->nctype,
                     cudaMemcpyHostToDevice, h->stream));
  SearchParams P;
  std::memset(&P, 0, sizeof(P));
  std::vector<double> lp;
  if (host) {
    lp.resize(npts);
    for (int i = 0; i < npts; ++i) lp[i] = std::log(rlev[i]);   // host libm
    for (int b = 0; b < 4; ++b) CK(h->cb[b].ensure(npts));
    CK(cudaMemcpyAsync(h->cb[0].p, ri, sizeof(double) * npts, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->cb[1].p, rj, sizeof(double) * npts, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->cb[2].p, lp.data(), sizeof(double) * npts, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->cb[3].p, rz, sizeof(double) * npts, cudaMemcpyHostToDevice, h->stream));
    P.ri = h->cb[0].p; P.rj = h->cb[1].p; P.lp = h->cb[2].p; P.rz = h->cb[3].p;
    CK(h->cb_i.ensure((size_t)npts * (1 + (idx ? max_out : 0))));
    P.nobsl = h->cb_i.p;
    P.idx = idx ? h->cb_i.p + npts : nullptr;
    if (rdiag) { CK(h->cb[4].ensure((size_t)npts * max_out)); P.rdiag = h->cb[4].p; }
    if (rloc) { CK(h->cb[5].ensure((size_t)npts * max_out)); P.rloc = h->cb[5].p; }
  } else {
    return fail(h, LETKF_B200_EINVAL, "obs_local: device mem_space not supported (host computes log p)");
  }
  const int grid = std::max(1, std::min(npts, h->num_sms * 8));
  CK(h->l_iob.ensure((size_t)grid * h->maxl));
  CK(h->l_rdiag.ensure((size_t)grid * h->maxl));
  CK(h->l_rloc.ensure((size_t)grid * h->maxl));
  CK(h->l_cnd.ensure((size_t)grid * h->ccap));
  CK(h->l_cpk.ensure((size_t)grid * h->ccap));
  P.l_cnd = h->l_cnd.p; P.l_cpk = h->l_cpk.p; P.ccap = h->ccap;
  CK(cudaMemsetAsync(h->counters.p, 0, 16 * sizeof(unsigned long long), h->stream));
  P.T = h->d_tables.p; P.rec = h->rec.p; P.bstart = h->bstart.p; P.vlfac = h->vlfac_one.p;
  P.npts = npts; P.max_out = max_out;
  P.l_iob = h->l_iob.p; P.l_rdiag = h->l_rdiag.p; P.l_rloc = h->l_rloc.p; P.lcap = h->maxl;
  P.counters = h->counters.p;
  // unused tail entries: idx = -1, rdiag = rloc = 0
  if (P.idx) CK(cudaMemsetAsync(P.idx, 0xFF, sizeof(int) * (size_t)npts * max_out, h->stream));
  if (P.rdiag) CK(cudaMemsetAsync(P.rdiag, 0, sizeof(double) * (size_t)npts * max_out, h->stream));
  if (P.rloc) CK(cudaMemsetAsync(P.rloc, 0, sizeof(double) * (size_t)npts * max_out, h->stream));
  if (npts > 0) {
    search_kernel<<<grid, 128, 0, h->stream>>>(P);
    CK(cudaGetLastError());
  }
  CK(cudaMemcpyAsync(nobsl, P.nobsl, sizeof(int) * npts, cudaMemcpyDeviceToHost, h->stream));
  if (idx) CK(cudaMemcpyAsync(idx, P.idx, sizeof(int) * (size_t)npts * max_out, cudaMemcpyDeviceToHost, h->stream));
  if (rdiag) CK(cudaMemcpyAsync(rdiag, P.rdiag, sizeof(double) * (size_t)npts * max_out, cudaMemcpyDeviceToHost, h->stream));
  if (rloc) CK(cudaMemcpyAsync(rloc, P.rloc, sizeof(double) * (size_t)npts * max_out, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (idx)
    for (int i = 0; i < npts; ++i)
      if (nobsl[i] > max_out) return fail(h, LETKF_B200_EINVAL, "obs_local: max_out too small");
  return LETKF_B200_OK;
}

int letkf_b200_nobs_out(letkf_b200_handle *h, int nvar, const double *pmean, const double *logp, double *out, int mem_space) {
  if (!h || !out || (!pmean && !logp)) return LETKF_B200_EINVAL;
  if (!h->obs_set) return fail(h, LETKF_B200_ESTATE, "set_obs has not been called");
  if (h->nij1 < 1) return fail(h, LETKF_B200_ESTATE, "set_grid has not been called");
  const letkf_b200_config &c = h->cfg;
  if (nvar < 1 || nvar > c.nv3d) return fail(h, LETKF_B200_EINVAL, "nobs_out: nvar must be a 3-D variable (1..nv3d)");
  CK(cudaSetDevice(h->device));
  const bool host = mem_space != LETKF_B200_MEM_DEVICE;
  const size_t sl = (size_t)h->nij1 * c.nlev;
  const int nct = std::max(h->nctype, 1);
  CK(h->vlfac_one.ensure(nct));
  CK(cudaMemcpyAsync(h->vlfac_one.p, h->h_vlfac.data() + (size_t)(nvar - 1) * h->nctype, sizeof(double) * h->nctype,
                     cudaMemcpyHostToDevice, h->stream));
  NobsOutParams P;
  std::memset(&P, 0, sizeof(P));
  std::vector<double> lp;
  if (host) {   // ln p by the host libm (selection thresholds bit-identical to a CPU run), like obs_local
    CK(h->st_logp.ensure(sl));
    const double *src = logp;
    if (!src) {
      lp.resize(sl);
      for (size_t i = 0; i < sl; ++i) lp[i] = std::log(pmean[i]);
      src = lp.data();
    }
    CK(cudaMemcpyAsync(h->st_logp.p, src, sizeof(double) * sl, cudaMemcpyHostToDevice, h->stream));
    P.logp = h->st_logp.p;
    CK(h->cb[4].ensure(sl * 11));
    P.out = h->cb[4].p;
  } else {
    P.logp = logp;
    P.pmean = pmean;
    P.out = out;
  }
  const int grid = (int)std::max<size_t>(1, std::min<size_t>(sl, (size_t)h->num_sms * 8));
  CK(h->l_iob.ensure((size_t)grid * h->maxl));
  CK(h->l_rdiag.ensure((size_t)grid * h->maxl));
  CK(h->l_rloc.ensure((size_t)grid * h->maxl));
  CK(h->l_cnd.ensure((size_t)grid * h->ccap));
  CK(h->l_cpk.ensure((size_t)grid * h->ccap));
  CK(cudaMemsetAsync(h->counters.p, 0, 16 * sizeof(unsigned long long), h->stream));
  P.T = h->d_tables.p; P.rec = h->rec.p; P.bstart = h->bstart.p; P.vlfac = h->vlfac_one.p;
  P.rig1 = h->rig1.p; P.rjg1 = h->rjg1.p; P.hgt1 = h->hgt1.p;
  P.nij1 = h->nij1; P.nlev = c.nlev;
  P.radar_only = h->radar_only ? 1 : 0;
  P.zcut = c.RADAR_ZMAX + std::max(c.VERT_LOCAL[21], c.VERT_LOCAL_RADAR_VR) * c.dist_zero_fac;
  P.BOUNDARY_BUFFER_WIDTH = c.BOUNDARY_BUFFER_WIDTH; P.DX = c.DX; P.DY = c.DY;
  P.IHALO = c.IHALO; P.JHALO = c.JHALO; P.nlon = c.nlon; P.nlat = c.nlat;
  P.l_iob = h->l_iob.p; P.l_rdiag = h->l_rdiag.p; P.l_rloc = h->l_rloc.p; P.lcap = h->maxl;
  P.l_cnd = h->l_cnd.p; P.l_cpk = h->l_cpk.p; P.ccap = h->ccap;
  P.counters = h->counters.p;
  nobs_out_kernel<<<grid, 128, 0, h->stream>>>(P);
  CK(cudaGetLastError());
  unsigned long long cnt[16];
  CK(cudaMemcpyAsync(cnt, h->counters.p, sizeof(cnt), cudaMemcpyDeviceToHost, h->stream));
  if (host) CK(cudaMemcpyAsync(out, P.out, sizeof(double) * sl * 11, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (cnt[5] != 0) return fail(h, LETKF_B200_EINVAL, "nobs_out: a local list exceeded the capacity sized by set_obs");
  return LETKF_B200_OK;
}

int letkf_b200_das_letkf(letkf_b200_handle *h, const letkf_b200_das_args *a) {
  if (!h || !a || !a->gues3d || !a->anal3d) return LETKF_B200_EINVAL;
  if (!h->obs_set) return fail(h, LETKF_B200_ESTATE, "set_obs has not been called");
  if (h->nij1 < 1) return fail(h, LETKF_B200_ESTATE, "set_grid has not been called");
  const letkf_b200_config &c = h->cfg;
  if (c.nv2d > 0 && (!a->gues2d || !a->anal2d)) return fail(h, LETKF_B200_EINVAL, "gues2d/anal2d required when nv2d > 0");
  if (c.INFL_MUL <= 0.0 && !a->infl3d) return fail(h, LETKF_B200_EINVAL, "INFL_MUL <= 0 needs the infl3d field");
  // 2-D variables take INFL_MUL: the reference's work2d field (read from the inflation file / adapted, letkf_tools.f90:
  // 240-262, 546) has no counterpart in this interface yet
  if (c.nv2d > 0 && (c.INFL_MUL <= 0.0 || c.INFL_MUL_ADAPTIVE))
    return fail(h, LETKF_B200_EINVAL, "nv2d > 0 with INFL_MUL <= 0 or INFL_MUL_ADAPTIVE is not supported (no 2-D inflation field)");
  CK(cudaSetDevice(h->device));
  const int k = c.MEMBER, nens = c.DET_RUN ? k + 2 : k + 1;
  const size_t sl = (size_t)h->nij1 * c.nlev;
  const size_t n3 = sl * nens * c.nv3d, n2 = (size_t)h->nij1 * nens * c.nv2d, nf = sl * c.nv3d;
  const bool host = a->mem_space != LETKF_B200_MEM_DEVICE;
  const bool back = !(a->reserved & 1);   // hand the destroyed gues (perturbations, mean) back to the host
  DasParams P = {};
  std::memset(&P, 0, sizeof(P));
  if (host) {
    CK(h->st_gues.ensure(n3));
    CK(h->st_anal.ensure(n3));
    P.gues3d = h->st_gues.p;
    P.anal3d = h->st_anal.p;
    ensure_pinned(h, a->gues3d, sizeof(double) * n3);
    ensure_pinned(h, a->anal3d, sizeof(double) * n3);
    if (c.nv2d > 0) {
      CK(h->st_gues2.ensure(n2));
      CK(h->st_anal2.ensure(n2));
      CK(cudaMemcpyAsync(h->st_gues2.p, a->gues2d, sizeof(double) * n2, cudaMemcpyHostToDevice, h->s_h2d));
      P.gues2d = h->st_gues2.p;
      P.anal2d = h->st_anal2.p;
    }
    if (a->infl3d) {
      CK(h->st_infl.ensure(nf));
      if (c.INFL_MUL <= 0.0) CK(cudaMemcpyAsync(h->st_infl.p, a->infl3d, sizeof(double) * nf, cudaMemcpyHostToDevice, h->s_h2d));
      P.infl3d = h->st_infl.p;
    }
    if (a->rtps_infl_out) { CK(h->st_rtps.ensure(nf)); P.rtps_out = h->st_rtps.p; }
    if (a->nobsl_out) { CK(h->st_nobsl.ensure(sl)); P.nobsl_out = h->st_nobsl.p; }
    {
      // ln(mean pressure) of every point on the HOST (letkf_tools.f90:1852-1866 takes log(rlev) with the host's libm):
      // the vertical-distance tests then see bit-for-bit what a CPU run on this machine sees.  A caller-supplied
      // table wins; with device-resident state and no table the kernels call CUDA's log() (<= 1 ulp).
      const double *lp = a->logp;
      if (!lp) {
        h->h_logp.resize(sl);
        const double *pm = a->gues3d + ((size_t)k + (size_t)(c.iv3d_p - 1) * nens) * sl;   // slot mmean of iv3d_p
        const unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        std::vector<std::thread> th;
        double *out = h->h_logp.data();
        for (unsigned t = 0; t < nt; ++t)
          th.emplace_back([=]() {
            for (size_t i = sl * t / nt; i < sl * (t + 1) / nt; ++i) out[i] = std::log(pm[i]);
          });
        for (auto &x : th) x.join();
        lp = out;
      }
      CK(h->st_logp.ensure(sl));
      CK(cudaMemcpyAsync(h->st_logp.p, lp, sizeof(double) * sl, cudaMemcpyHostToDevice, h->s_h2d));
      P.logp = h->st_logp.p;
    }
  } else {
    P.gues3d = a->gues3d; P.anal3d = a->anal3d; P.gues2d = a->gues2d; P.anal2d = a->anal2d;
    P.infl3d = a->infl3d; P.rtps_out = a->rtps_infl_out; P.nobsl_out = a->nobsl_out; P.logp = a->logp;
  }
  if (P.infl3d) {   // work3d = INFL_MUL (or the field read by the caller), clamped from below (letkf_tools.f90:240-267):
                    // every point carries its value on return, also the ones the analysis skips
    cudaStream_t sf = host ? h->s_h2d : h->stream;   // host mode: behind the upload of the field, before ev_in[0]
    if (c.INFL_MUL > 0.0) fill_kernel<<<h->num_sms * 8, 256, 0, sf>>>(P.infl3d, nf, c.INFL_MUL);
    if (c.INFL_MUL_MIN > 0.0) clamp_min_kernel<<<h->num_sms * 8, 256, 0, sf>>>(P.infl3d, nf, c.INFL_MUL_MIN);
    CK(cudaGetLastError());
  }
  if (P.rtps_out) {   // work3da = 1 (letkf_tools.f90:271-276); skipped points keep 1
    fill_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(P.rtps_out, nf, 1.0);
    CK(cudaGetLastError());
  }
  if (P.nobsl_out) CK(cudaMemsetAsync(P.nobsl_out, 0, sizeof(int) * sl, h->stream));
  P.k = k; P.nens = nens; P.nij1 = h->nij1; P.nlev = c.nlev; P.nv3d = c.nv3d; P.nv2d = c.nv2d; P.det = c.DET_RUN ? 1 : 0;
  P.ld = ld_of(k); P.ldk = ldk_of(k); P.npairs = (k + 1) / 2; P.ncols = 2 * P.npairs;
  P.rig1 = h->rig1.p; P.rjg1 = h->rjg1.p; P.hgt1 = h->hgt1.p;
  P.T = h->d_tables.p; P.rec = h->rec.p; P.bstart = h->bstart.p; P.ensval = h->sens.p; P.val = h->sval.p; P.ldens = h->ldens;
  P.nvgroup = h->nvgroup;
  for (int n = 0; n < kMaxNV; ++n) { P.vgroup[n] = h->vgroup[n]; P.vfirst[n] = h->vfirst[n]; P.gmask[n] = 0; }
  for (int n = 0; n < c.nv3d + c.nv2d; ++n) P.gmask[h->vgroup[n]] |= 1u << n;
  P.qmask = 0;
  for (int n = 0; n < c.nv3d; ++n)
    if (n + 1 >= c.iv3d_q && n + 1 <= c.iv3d_qg) P.qmask |= 1u << n;
  P.vlfac = h->vlfac_groups.p;
  P.INFL_MUL = c.INFL_MUL; P.INFL_MUL_MIN = c.INFL_MUL_MIN; P.RELAX_ALPHA = c.RELAX_ALPHA;
  P.RELAX_ALPHA_SPREAD = c.RELAX_ALPHA_SPREAD; P.Q_UPDATE_TOP = c.Q_UPDATE_TOP; P.Q_SPRD_MAX = c.Q_SPRD_MAX;
  P.RELAX_TO_INFLATED_PRIOR = c.RELAX_TO_INFLATED_PRIOR; P.INFL_MUL_ADAPTIVE = c.INFL_MUL_ADAPTIVE;
  P.infl_from_field = (c.INFL_MUL <= 0.0) ? 1 : 0;
  P.iv3d_p = c.iv3d_p; P.iv3d_q = c.iv3d_q; P.iv3d_qg = c.iv3d_qg;
  P.radar_only = h->radar_only ? 1 : 0;
  P.zcut = c.RADAR_ZMAX + std::max(c.VERT_LOCAL[21], c.VERT_LOCAL_RADAR_VR) * c.dist_zero_fac;
  P.BOUNDARY_BUFFER_WIDTH = c.BOUNDARY_BUFFER_WIDTH; P.DX = c.DX; P.DY = c.DY;
  P.IHALO = c.IHALO; P.JHALO = c.JHALO; P.nlon = c.nlon; P.nlat = c.nlat;
  P.lcap = h->maxl;
  P.counters = h->counters.p;
  P.max_sweeps = 30;
  P.stagger_ns = 0;
  P.stagger_div = std::max(h->num_sms, 1);
  if (const char *sg = std::getenv("LETKF_B200_STAGGER_US")) P.stagger_ns = (int)(1000.0 * std::atof(sg));
  CK(cudaMemsetAsync(h->counters.p, 0, 16 * sizeof(unsigned long long), h->stream));
#ifdef LETKF_EXP_TRACE
  CK(h->cb[9].ensure(32000));
  CK(cudaMemsetAsync(h->cb[9].p, 0, sizeof(double) * 32000, h->stream));
  P.trace = reinterpret_cast<long long *>(h->cb[9].p);
#endif
  int r, nlaunch = 0;
  DasLaunch L;
  // MEMBER <= 102: tensor-core Newton-Schulz solve, one CTA per point (das_ns_kernel.cuh); larger
  // ensembles (the k x k matrices no longer fit in shared memory): tiled whole-GPU path (tiled.cuh).
  // LETKF_B200_SOLVER=jacobi (MEMBER <= 128): Cholesky + one-sided Jacobi; =tiled forces the tiled path.
  const char *sv = std::getenv("LETKF_B200_SOLVER");
  const int nsc = ns_class(k);
  const bool jacobi = sv && std::strcmp(sv, "jacobi") == 0 && k <= 128;
  const bool tiled = !jacobi && (nsc == 0 || (sv && std::strcmp(sv, "tiled") == 0));
  if (tiled) {
    if (!h->tiled) h->tiled = new TiledBufs();
    r = LETKF_B200_OK;
  } else if (!jacobi) {
    r = plan_das_ns_class<false>(h, nsc, P, L);
  } else
  if (k <= 20) r = plan_das<20>(h, P, L);
  else if (k <= 52) r = plan_das<52>(h, P, L);
  else if (k <= 64) r = plan_das<64>(h, P, L);
  else if (k <= 100) r = plan_das<100>(h, P, L);
  else r = plan_das<128>(h, P, L);
  if (r != LETKF_B200_OK) return r;

  // Level chunks.  Device-resident state: 12 chunks (pre-search pipeline).  Host state: 24 chunks pipelined over three
  // streams -- H2D of chunk c+1 and D2H of chunk c-1 overlap the analysis of chunk c (PCIe is full
  // duplex), so the call costs max(copy, compute) instead of their sum.
  // Pre-search (one-CTA-per-point solver only): the local-observation search of chunk c+1 runs in its own
  // kernel on a high-priority stream while the solver works on chunk c, so the latency-bound search fills
  // the issue slots the tensor-core solver leaves idle.  LETKF_B200_PRESEARCH=0 searches inside the solver.
  const char *pe = std::getenv("LETKF_B200_PRESEARCH");
  const bool pre = !tiled && !jacobi && !(pe && pe[0] == '0');
  DasLaunch Lpre;   // the search-free solver; L (with in-kernel search) then only runs the redo pass
  if (pre) {
    r = plan_das_ns_class<true>(h, nsc, P, Lpre);
    if (r != LETKF_B200_OK) return r;
  }
  int nchunk = 1;
  {
    const char *ce = std::getenv("LETKF_B200_CHUNKS");
    if (host) nchunk = ce ? std::atoi(ce) : 24;   // short pipeline ramps: C2 e2e 888 ms (10 chunks) -> 850 ms
    else if (pre) nchunk = ce ? std::atoi(ce) : 12;
    nchunk = std::max(1, std::min(nchunk, c.nlev));
  }
  while ((int)h->ev_in.size() < nchunk) {
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); h->ev_in.push_back(e);
    CK(cudaEventCreate(&e)); h->ev_k0.push_back(e);
    CK(cudaEventCreate(&e)); h->ev_k1.push_back(e);
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); h->ev_s.push_back(e);
  }
  const size_t pitch = sizeof(double) * sl;   // distance between (member, variable) planes
  const size_t planes = (size_t)nens * c.nv3d;
  auto lev0 = [&](int ch) { return (int)((long long)c.nlev * ch / nchunk); };
  long long pl_cap = 0;
  size_t pl_entries_max = 0;
  if (pre) {
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, presearch_kernel, 128, 0));
    h->ps_grid = std::max(1, std::min(occ, 8)) * h->num_sms;
    CK(h->ps_l_iob.ensure((size_t)h->ps_grid * h->maxl)); CK(h->ps_l_rdiag.ensure((size_t)h->ps_grid * h->maxl));
    CK(h->ps_l_rloc.ensure((size_t)h->ps_grid * h->maxl));
    CK(h->ps_l_cnd.ensure((size_t)h->ps_grid * h->ccap)); CK(h->ps_l_cpk.ensure((size_t)h->ps_grid * h->ccap));
    int maxlev = 0;
    for (int ch = 0; ch < nchunk; ++ch) maxlev = std::max(maxlev, lev0(ch + 1) - lev0(ch));
    pl_entries_max = (size_t)maxlev * h->nij1 * h->nvgroup;
    // pool budget per buffer: what the longest chunk can need, within [2 GB, 8 GB] and a fifth of the free memory
    // (a pool that overflows sends points to the redo pass, whose in-kernel search is ~2x slower: C3 on one GPU)
    const char *pm = std::getenv("LETKF_B200_POOL_MB");
    double budget = 2048.0 * 1048576.0;
    if (pm) {
      budget = std::atof(pm) * 1048576.0;
    } else {
      size_t fr = 0, tot = 0;
      if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) {
        double held = 0.0;   // what the two pools already hold counts as free
        for (int b = 0; b < letkf_b200_handle::kPools; ++b) held += (double)h->pl_iob[b].n * 4.0 + (double)h->pl_rdiag[b].n * 8.0 + (double)h->pl_rloc[b].n * 8.0;
        budget = std::max(budget, std::min(8192.0 * 1048576.0, 0.2 * ((double)fr + held) / (double)letkf_b200_handle::kPools));
      }
    }
    const size_t per = c.INFL_MUL_ADAPTIVE ? 20 : 12;
    pl_cap = (long long)std::min<double>((double)pl_entries_max * (double)round_up(h->maxl, 4), budget / per);
    pl_cap = std::max<long long>(pl_cap, 4);
    CK(h->redo_list.ensure((size_t)maxlev * h->nij1));
    for (int b = 0; b < letkf_b200_handle::kPools; ++b) {
      CK(h->pl_n[b].ensure(pl_entries_max)); CK(h->pl_off[b].ensure(pl_entries_max));
      CK(h->pl_iob[b].ensure((size_t)pl_cap)); CK(h->pl_rdiag[b].ensure((size_t)pl_cap));
      if (c.INFL_MUL_ADAPTIVE) CK(h->pl_rloc[b].ensure((size_t)pl_cap));
    }
  }
  auto launch_search = [&](int ch) -> int {
    const int b = ch % letkf_b200_handle::kPools, l0 = lev0(ch), l1 = lev0(ch + 1);
    if (host) CK(cudaStreamWaitEvent(h->s_search, h->ev_in[ch], 0));
    if (ch >= letkf_b200_handle::kPools)   // the pool of chunk ch - 3 is free again
      CK(cudaStreamWaitEvent(h->s_search, h->ev_k1[ch - letkf_b200_handle::kPools], 0));
    DasParams Q = P;
    Q.point_begin = (long long)l0 * h->nij1;
    Q.point_end = (long long)l1 * h->nij1;
    Q.pl_base = Q.point_begin;
    Q.pl_cap = pl_cap;
    Q.pl_iob = h->pl_iob[b].p; Q.pl_rdiag = h->pl_rdiag[b].p; Q.pl_rloc = c.INFL_MUL_ADAPTIVE ? h->pl_rloc[b].p : nullptr;
    Q.pl_cursor = h->ps_counters.p + 20 + b;
    Q.counters = h->ps_counters.p;
    Q.l_iob = h->ps_l_iob.p; Q.l_rdiag = h->ps_l_rdiag.p; Q.l_rloc = h->ps_l_rloc.p; Q.l_cnd = h->ps_l_cnd.p; Q.l_cpk = h->ps_l_cpk.p;
    Q.ccap = h->ccap;
    CK(cudaMemsetAsync(h->ps_counters.p, 0, sizeof(unsigned long long), h->s_search));
    CK(cudaMemsetAsync(h->ps_counters.p + 20 + b, 0, sizeof(unsigned long long), h->s_search));
    const long long npts = Q.point_end - Q.point_begin;
    presearch_kernel<<<(unsigned)std::min<long long>(h->ps_grid, std::max<long long>(npts, 1)), 128, 0, h->s_search>>>(
        Q, h->pl_n[b].p, h->pl_off[b].p);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev_s[ch], h->s_search));
    return LETKF_B200_OK;
  };
  if (host) {
    for (int ch = 0; ch < nchunk; ++ch) {
      const int l0 = lev0(ch), l1 = lev0(ch + 1);
      const size_t off = (size_t)l0 * h->nij1, w = sizeof(double) * (size_t)(l1 - l0) * h->nij1;
      CK(cudaMemcpy2DAsync(P.gues3d + off, pitch, a->gues3d + off, pitch, w, planes, cudaMemcpyHostToDevice, h->s_h2d));
      CK(cudaEventRecord(h->ev_in[ch], h->s_h2d));
    }
  }
  if (pre) {
    // the search stream must not start before the work already queued on the compute stream (state restore,
    // obs tables) is done
    CK(cudaEventRecord(h->ev0, h->stream));
    CK(cudaStreamWaitEvent(h->s_search, h->ev0, 0));
    // The search runs TWO chunks ahead: the search of chunk c + 2 becomes eligible when the solver of chunk c - 1 is done, i.e.
    // together with the solver of chunk c -- whichever of the two gets the SMs first, the solver of chunk c + 1 finds its
    // lists ready.  (One chunk ahead, two pools: when the persistent solver CTAs of chunk c won that race the search of
    // chunk c + 1 only ran in their tail and the solver of chunk c + 1 waited for it: 1 - 5 ms per chunk, box to box.)
    r = launch_search(0);
    if (r != LETKF_B200_OK) return r;
    if (nchunk > 1) {
      r = launch_search(1);
      if (r != LETKF_B200_OK) return r;
    }
  }
  for (int ch = 0; ch < nchunk; ++ch) {
    const int l0 = lev0(ch), l1 = lev0(ch + 1);
    if (pre && ch + 2 < nchunk) {
      r = launch_search(ch + 2);
      if (r != LETKF_B200_OK) return r;
    }
    if (host) CK(cudaStreamWaitEvent(h->stream, h->ev_in[ch], 0));
    if (pre) {
      CK(cudaStreamWaitEvent(h->stream, h->ev_s[ch], 0));
      const int b = ch % letkf_b200_handle::kPools;
      P.pl_n = h->pl_n[b].p; P.pl_off = h->pl_off[b].p; P.pl_iob = h->pl_iob[b].p; P.pl_rdiag = h->pl_rdiag[b].p;
      P.pl_rloc = c.INFL_MUL_ADAPTIVE ? h->pl_rloc[b].p : nullptr;
      P.pl_base = (long long)l0 * h->nij1;
    }
    if (tiled) {
      r = launch_range_tiled(h, *h->tiled, P, (long long)l0 * h->nij1, (long long)l1 * h->nij1, h->ev_k0[ch], h->ev_k1[ch], &nlaunch);
    } else if (pre) {
      // search-free solver over the chunk, then the redo pass (normally empty) with in-kernel search
      const long long pb = (long long)l0 * h->nij1, pe2 = (long long)l1 * h->nij1;
      P.redo_list = h->redo_list.p;
      P.redo_count = h->ps_counters.p + 18;
      P.point_list = nullptr;
      P.point_count = nullptr;
      CK(cudaMemsetAsync(h->ps_counters.p + 18, 0, sizeof(unsigned long long), h->stream));
      CK(cudaMemsetAsync(h->counters.p, 0, sizeof(unsigned long long), h->stream));
      CK(cudaEventRecord(h->ev_k0[ch], h->stream));
      {
        DasParams Q = P;
        Q.point_begin = pb;
        Q.point_end = pe2;
        Lpre.kern<<<(unsigned)std::min<long long>(Lpre.grid, std::max<long long>(pe2 - pb, 1)), Lpre.nt, Lpre.smem, h->stream>>>(Q);
        CK(cudaGetLastError());
        Q.pl_n = nullptr;            // redo pass: the solver searches by itself
        Q.pl_off = nullptr;
        Q.point_list = h->redo_list.p;
        Q.point_count = h->ps_counters.p + 18;
        Q.point_begin = 0;
        Q.point_end = pe2 - pb;      // upper bound; the real length is *point_count
        CK(cudaMemsetAsync(h->counters.p, 0, sizeof(unsigned long long), h->stream));
              L.kern<<<(unsigned)std::min<long long>(L.grid, std::max<long long>(pe2 - pb, 1)), L.nt, L.smem, h->stream>>>(Q);
        CK(cudaGetLastError());
      }
      CK(cudaEventRecord(h->ev_k1[ch], h->stream));
      r = LETKF_B200_OK;
    } else {
      r = launch_range(h, L, P, (long long)l0 * h->nij1, (long long)l1 * h->nij1, h->ev_k0[ch], h->ev_k1[ch]);
    }
    if (r != LETKF_B200_OK) return r;
    if (host) {
      const size_t off = (size_t)l0 * h->nij1, w = sizeof(double) * (size_t)(l1 - l0) * h->nij1;
      CK(cudaStreamWaitEvent(h->s_d2h, h->ev_k1[ch], 0));
      CK(cudaMemcpy2DAsync(a->anal3d + off, pitch, P.anal3d + off, pitch, w, planes, cudaMemcpyDeviceToHost, h->s_d2h));
      // gues3d is INTENT(INOUT) "destroyed" in the reference; hand the perturbations back too so
      // that callers relying on slots 1..k holding dX / slot k+1 the mean keep working.
      if (back)
        CK(cudaMemcpy2DAsync(a->gues3d + off, pitch, P.gues3d + off, pitch, w, planes, cudaMemcpyDeviceToHost, h->s_d2h));
    }
  }
  h->last_launches = tiled ? nlaunch : (pre ? 3 * nchunk : nchunk);   // pre: presearch + solver + redo pass per chunk
  if (host) {
    if (c.nv2d > 0) {
      CK(cudaMemcpyAsync(a->anal2d, P.anal2d, sizeof(double) * n2, cudaMemcpyDeviceToHost, h->s_d2h));
      if (back) CK(cudaMemcpyAsync(a->gues2d, P.gues2d, sizeof(double) * n2, cudaMemcpyDeviceToHost, h->s_d2h));
    }
    if (a->infl3d) CK(cudaMemcpyAsync(a->infl3d, P.infl3d, sizeof(double) * nf, cudaMemcpyDeviceToHost, h->s_d2h));
    if (a->rtps_infl_out) CK(cudaMemcpyAsync(a->rtps_infl_out, P.rtps_out, sizeof(double) * nf, cudaMemcpyDeviceToHost, h->s_d2h));
    if (a->nobsl_out) CK(cudaMemcpyAsync(a->nobsl_out, P.nobsl_out, sizeof(int) * sl, cudaMemcpyDeviceToHost, h->s_d2h));
  }
  unsigned long long cnt[16];
  CK(cudaMemcpyAsync(cnt, h->counters.p, sizeof(cnt), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (host) CK(cudaStreamSynchronize(h->s_d2h));
#ifdef LETKF_EXP_TRACE
  if (const char *tp = std::getenv("LETKF_B200_TRACE")) {
    std::vector<long long> tr(32000);
    cudaMemcpy(tr.data(), h->cb[9].p, sizeof(long long) * 32000, cudaMemcpyDeviceToHost);
    if (FILE *f = std::fopen(tp, "w")) {
      for (int i = 0; i < 16000 && tr[2 * i] != 0; ++i) std::fprintf(f, "%lld %lld\n", tr[2 * i], tr[2 * i + 1]);
      std::fclose(f);
    }
  }
#endif
  h->last_ms = 0.f;
  for (int ch = 0; ch < nchunk; ++ch) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev_k0[ch], h->ev_k1[ch]));
    h->last_ms += ms;
  }
  h->st_sweeps = (long long)cnt[6];
  h->st_refined = (long long)cnt[7];
  for (int i = 0; i < 8; ++i) h->st_phase[i] = (long long)cnt[8 + i];
  h->st_points = (long long)cnt[1];
  h->st_solved = (long long)cnt[2];
  h->st_fail = (long long)cnt[3];
  h->st_nobs = (long long)cnt[4];
  if (cnt[5] > 0) return fail(h, LETKF_B200_ENOMEM, "local observation list overflow");
  if (cnt[3] > 0) return fail(h, LETKF_B200_EEIGEN, "eigensolve failed at one or more grid points");
  return LETKF_B200_OK;
}

int letkf_b200_das_stats(const letkf_b200_handle *h, int64_t *npoints, int64_t *nsolved, int64_t *nfail,
                         int64_t *nobsl_sum) {
  if (!h) return LETKF_B200_EINVAL;
  if (npoints) *npoints = h->st_points;
  if (nsolved) *nsolved = h->st_solved;
  if (nfail) *nfail = h->st_fail;
  if (nobsl_sum) *nobsl_sum = h->st_nobs;
  return LETKF_B200_OK;
}
int letkf_b200_das_refined(const letkf_b200_handle *h, int64_t *nrefined) {
  if (!h || !nrefined) return LETKF_B200_EINVAL;
  *nrefined = h->st_refined;
  return LETKF_B200_OK;
}
int letkf_b200_das_kernel_ms(const letkf_b200_handle *h, float *ms, int *launches) {
  if (!h) return LETKF_B200_EINVAL;
  if (ms) *ms = h->last_ms;
  if (launches) *launches = h->last_launches;
  return LETKF_B200_OK;
}

int letkf_b200_das_phase_clocks(const letkf_b200_handle *h, int64_t *clocks, int64_t *solver_iterations) {
  if (!h) return LETKF_B200_EINVAL;
  if (clocks)
    for (int i = 0; i < 8; ++i) clocks[i] = h->st_phase[i];
  if (solver_iterations) *solver_iterations = h->st_sweeps;
  return LETKF_B200_OK;
}

int letkf_b200_core_batch(letkf_b200_handle *h, int ne, int nobs, int npts, const int32_t *nobsl, const double *hdxb,
                          const double *rdiag, const double *rloc, const double *dep, double *parm_infl, double *trans,
                          double *transm, double *pao, int rdiag_wloc, int infl_update, const double *depd,
                          double *transmd, int mem_space) {
  if (!h || ne < 2 || ne > LETKF_B200_MAX_MEMBER || nobs < 0 || npts < 0) return LETKF_B200_EINVAL;
  if (!nobsl || !hdxb || !rdiag || !rloc || !dep || !parm_infl || !trans) return LETKF_B200_EINVAL;
  CK(cudaSetDevice(h->device));
  const bool host = mem_space != LETKF_B200_MEM_DEVICE;
  const size_t k2 = (size_t)ne * ne, no = (size_t)npts * nobs;
  CoreParams P;
  std::memset(&P, 0, sizeof(P));
  if (host) {
    for (int i = 0; i < npts; ++i)
      if (nobsl[i] < 0 || nobsl[i] > nobs) return fail(h, LETKF_B200_EINVAL, "nobsl out of range");
    CK(h->cb_i.ensure(npts));
    CK(h->cb[0].ensure(no * ne)); CK(h->cb[1].ensure(no)); CK(h->cb[2].ensure(no)); CK(h->cb[3].ensure(no));
    CK(h->cb[4].ensure(npts)); CK(h->cb[5].ensure(npts * k2));
    CK(cudaMemcpyAsync(h->cb_i.p, nobsl, sizeof(int) * npts, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->cb[0].p, hdxb, sizeof(double) * no * ne, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->cb[1].p, rdiag, sizeof(double) * no, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->cb[2].p, rloc, sizeof(double) * no, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->cb[3].p, dep, sizeof(double) * no, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->cb[4].p, parm_infl, sizeof(double) * npts, cudaMemcpyHostToDevice, h->stream));
    P.nobsl = h->cb_i.p; P.hdxb = h->cb[0].p; P.rdiag = h->cb[1].p; P.rloc = h->cb[2].p; P.dep = h->cb[3].p;
    P.parm_infl = h->cb[4].p; P.trans = h->cb[5].p;
    if (transm) { CK(h->cb[6].ensure((size_t)npts * ne)); P.transm = h->cb[6].p; }
    if (pao) { CK(h->cb[7].ensure(npts * k2)); P.pao = h->cb[7].p; }
    if (depd) {
      CK(h->cb[8].ensure(no));
      CK(cudaMemcpyAsync(h->cb[8].p, depd, sizeof(double) * no, cudaMemcpyHostToDevice, h->stream));
      P.depd = h->cb[8].p;
    }
    if (transmd) { CK(h->cb[9].ensure((size_t)npts * ne)); P.transmd = h->cb[9].p; }
  } else {
    P.nobsl = nobsl; P.hdxb = hdxb; P.rdiag = rdiag; P.rloc = rloc; P.dep = dep; P.parm_infl = parm_infl;
    P.trans = trans; P.transm = transm; P.pao = pao; P.depd = depd; P.transmd = transmd;
  }
  P.ne = ne; P.nobs = nobs; P.npts = npts;
  P.ld = ld_of(ne); P.ldk = ldk_of(ne); P.npairs = (ne + 1) / 2; P.ncols = 2 * P.npairs;
  P.rdiag_wloc = rdiag_wloc; P.infl_update = infl_update;
  P.counters = h->counters.p;
  P.max_sweeps = 30;
  CK(cudaMemsetAsync(h->counters.p, 0, 16 * sizeof(unsigned long long), h->stream));
  if (npts > 0 && ne > 128) {   // large ensembles: tiled whole-GPU solve (tiled.cuh)
    if (!h->tiled) h->tiled = new TiledBufs();
    CoreTiledParams C;
    std::memset(&C, 0, sizeof(C));
    C.ne = ne; C.nobs = nobs; C.npts = npts; C.rdiag_wloc = rdiag_wloc; C.infl_update = infl_update;
    C.nobsl = P.nobsl; C.hdxb = P.hdxb; C.rdiag = P.rdiag; C.rloc = P.rloc; C.dep = P.dep; C.depd = P.depd;
    C.parm_infl = P.parm_infl; C.trans = P.trans; C.transm = P.transm; C.pao = P.pao; C.transmd = P.transmd;
    C.counters = h->counters.p;
    const int r = core_batch_tiled(h, *h->tiled, C);
    if (r != LETKF_B200_OK) return r;
  } else if (npts > 0) {
    int r;
    if (ne <= 20) r = launch_core<20>(h, P);
    else if (ne <= 52) r = launch_core<52>(h, P);
    else if (ne <= 64) r = launch_core<64>(h, P);
    else if (ne <= 100) r = launch_core<100>(h, P);
    else r = launch_core<128>(h, P);
    if (r != LETKF_B200_OK) return r;
  }
  if (host) {
    CK(cudaMemcpyAsync(trans, P.trans, sizeof(double) * npts * k2, cudaMemcpyDeviceToHost, h->stream));
    if (transm) CK(cudaMemcpyAsync(transm, P.transm, sizeof(double) * (size_t)npts * ne, cudaMemcpyDeviceToHost, h->stream));
    if (pao) CK(cudaMemcpyAsync(pao, P.pao, sizeof(double) * npts * k2, cudaMemcpyDeviceToHost, h->stream));
    if (transmd) CK(cudaMemcpyAsync(transmd, P.transmd, sizeof(double) * (size_t)npts * ne, cudaMemcpyDeviceToHost, h->stream));
    if (infl_update) CK(cudaMemcpyAsync(parm_infl, P.parm_infl, sizeof(double) * npts, cudaMemcpyDeviceToHost, h->stream));
  }
  unsigned long long cnt[8];
  CK(cudaMemcpyAsync(cnt, h->counters.p, sizeof(cnt), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (cnt[3] > 0) return fail(h, LETKF_B200_EEIGEN, "eigensolve failed at one or more points");
  return LETKF_B200_OK;
}

int letkf_b200_ensmean_grd(letkf_b200_handle *h, int mem, int nens, int nij, double *v3d, double *v2d, int mem_space) {
  if (!h || mem < 1 || nens <= mem || nij < 1 || !v3d) return LETKF_B200_EINVAL;
  CK(cudaSetDevice(h->device));
  const letkf_b200_config &c = h->cfg;
  const size_t sl = (size_t)nij * c.nlev, n3 = sl * nens * c.nv3d, n2 = (size_t)nij * nens * c.nv2d;
  double *d3 = v3d, *d2 = v2d;
  if (mem_space != LETKF_B200_MEM_DEVICE) {
    CK(h->st_gues.ensure(n3));
    CK(cudaMemcpyAsync(h->st_gues.p, v3d, sizeof(double) * n3, cudaMemcpyHostToDevice, h->stream));
    d3 = h->st_gues.p;
    if (c.nv2d > 0 && v2d) {
      CK(h->st_gues2.ensure(n2));
      CK(cudaMemcpyAsync(h->st_gues2.p, v2d, sizeof(double) * n2, cudaMemcpyHostToDevice, h->stream));
      d2 = h->st_gues2.p;
    }
  }
  {
    const size_t tot = sl * c.nv3d;
    ensmean_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(mem, nens, sl, c.nv3d, d3);
  }
  if (c.nv2d > 0 && d2) {
    const size_t tot = (size_t)nij * c.nv2d;
    ensmean_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(mem, nens, (size_t)nij, c.nv2d, d2);
  }
  CK(cudaGetLastError());
  if (mem_space != LETKF_B200_MEM_DEVICE) {
    CK(cudaMemcpyAsync(v3d, d3, sizeof(double) * n3, cudaMemcpyDeviceToHost, h->stream));
    if (c.nv2d > 0 && v2d) CK(cudaMemcpyAsync(v2d, d2, sizeof(double) * n2, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return LETKF_B200_OK;
}

int letkf_b200_enssprd_grd(letkf_b200_handle *h, int mem, int nens, int nij, const double *v3d, const double *v2d,
                           double *v3ds, double *v2ds, int mem_space) {
  if (!h || mem < 2 || nens <= mem || nij < 1 || !v3d || !v3ds) return LETKF_B200_EINVAL;
  CK(cudaSetDevice(h->device));
  const letkf_b200_config &c = h->cfg;
  const size_t sl = (size_t)nij * c.nlev, n3 = sl * nens * c.nv3d, n2 = (size_t)nij * nens * c.nv2d;
  const size_t o3 = sl * c.nv3d, o2 = (size_t)nij * c.nv2d;
  const bool host = mem_space != LETKF_B200_MEM_DEVICE, two = c.nv2d > 0 && v2d && v2ds;
  const double *d3 = v3d, *d2 = v2d;
  double *s3 = v3ds, *s2 = v2ds;
  if (host) {
    CK(h->st_gues.ensure(n3)); CK(h->st_rtps.ensure(o3));
    CK(cudaMemcpyAsync(h->st_gues.p, v3d, sizeof(double) * n3, cudaMemcpyHostToDevice, h->stream));
    d3 = h->st_gues.p; s3 = h->st_rtps.p;
    if (two) {
      CK(h->st_gues2.ensure(n2)); CK(h->st_anal2.ensure(o2));
      CK(cudaMemcpyAsync(h->st_gues2.p, v2d, sizeof(double) * n2, cudaMemcpyHostToDevice, h->stream));
      d2 = h->st_gues2.p; s2 = h->st_anal2.p;
    }
  }
  enssprd_kernel<<<(unsigned)((o3 + 255) / 256), 256, 0, h->stream>>>(mem, nens, sl, c.nv3d, d3, s3);
  if (two) enssprd_kernel<<<(unsigned)((o2 + 255) / 256), 256, 0, h->stream>>>(mem, nens, (size_t)nij, c.nv2d, d2, s2);
  CK(cudaGetLastError());
  if (host) {
    CK(cudaMemcpyAsync(v3ds, s3, sizeof(double) * o3, cudaMemcpyDeviceToHost, h->stream));
    if (two) CK(cudaMemcpyAsync(v2ds, s2, sizeof(double) * o2, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return LETKF_B200_OK;
}

int letkf_b200_additive_inflation(letkf_b200_handle *h, double infl_add, int q_ratio, int ref_only, const int32_t *ishuf,
                                  const double *addi3d, const double *addi2d, const double *gues3d, double *anal3d,
                                  double *anal2d, double *weight_out, int mem_space) {
  if (!h || !addi3d || !anal3d || !(infl_add > 0.0) || (q_ratio && !gues3d)) return LETKF_B200_EINVAL;
  if (h->nij1 < 1) return fail(h, LETKF_B200_ESTATE, "additive_inflation: set_grid has not been called");
  CK(cudaSetDevice(h->device));
  const letkf_b200_config &c = h->cfg;
  const int k = c.MEMBER, nens = c.DET_RUN ? k + 2 : k + 1, nij = h->nij1;
  const size_t sl = (size_t)nij * c.nlev, n3 = sl * nens * c.nv3d, n2 = (size_t)nij * nens * c.nv2d, o3 = sl * c.nv3d;
  const bool host = mem_space != LETKF_B200_MEM_DEVICE, two = c.nv2d > 0 && addi2d && anal2d;
  if (ishuf)
    for (int m = 0; m < k; ++m)
      if (ishuf[m] < 1 || ishuf[m] > k) return fail(h, LETKF_B200_EINVAL, "additive_inflation: ishuf is not a permutation of 1..MEMBER");
  // ---- addinfl_weight (:812-842) ----
  const double *d_w = nullptr;
  if (ref_only) {
    if (h->nobstotal < 0 || h->h_bstart.empty()) return fail(h, LETKF_B200_ESTATE, "additive_inflation: INFL_ADD_REF_ONLY needs set_obs");
    CK(h->st_logp.ensure((size_t)nij));
    int b0 = 0, b1 = 0;
    double hloc = 1.0;
    for (int ic = 0; ic < h->tables.nctype; ++ic) {   // ctype_elmtyp(uid_obs(id_radar_ref_obs), 22)
      const CtypeDev &d = h->tables.ct[ic];
      if (d.elm_u == uid_obs(ID_REF) && d.typ == 22) {
        b0 = h->h_bstart[d.boff];
        b1 = h->h_bstart[d.boff + d.ngrdext_i * d.ngrdext_j];
        hloc = d.hori_loc;
      }
    }
    addinfl_weight_kernel<<<(nij + 255) / 256, 256, 0, h->stream>>>(nij, h->rig1.p, h->rjg1.p, h->rec.p, b0, b1, c.DX, c.DY, hloc,
                                                                   c.dist_zero_fac_square, h->st_logp.p);
    CK(cudaGetLastError());
    d_w = h->st_logp.p;
    if (weight_out) {
      CK(cudaMemcpyAsync(weight_out, d_w, sizeof(double) * nij, host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, h->stream));
    }
  } else if (weight_out) {
    std::vector<double> ones((size_t)nij, 1.0);
    CK(cudaMemcpyAsync(weight_out, ones.data(), sizeof(double) * nij, host ? cudaMemcpyHostToHost : cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  // ---- the update (:869-925) ----
  const double *d_add3 = addi3d, *d_add2 = addi2d;
  double *d_an3 = anal3d, *d_an2 = anal2d;
  if (host) {
    CK(h->st_gues.ensure(n3)); CK(h->st_anal.ensure(n3));
    CK(cudaMemcpyAsync(h->st_gues.p, addi3d, sizeof(double) * n3, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->st_anal.p, anal3d, sizeof(double) * n3, cudaMemcpyHostToDevice, h->stream));
    d_add3 = h->st_gues.p; d_an3 = h->st_anal.p;
    if (q_ratio) {   // only the mean slot of the background is read: one plane per variable, at its place in a gues3d-shaped view
      CK(h->st_rtps.ensure(o3));
      CK(cudaMemcpy2DAsync(h->st_rtps.p, sizeof(double) * sl, gues3d + (size_t)k * sl, sizeof(double) * sl * nens, sizeof(double) * sl,
                           c.nv3d, cudaMemcpyHostToDevice, h->stream));
    }
    if (two) {
      CK(h->st_gues2.ensure(n2)); CK(h->st_anal2.ensure(n2));
      CK(cudaMemcpyAsync(h->st_gues2.p, addi2d, sizeof(double) * n2, cudaMemcpyHostToDevice, h->stream));
      CK(cudaMemcpyAsync(h->st_anal2.p, anal2d, sizeof(double) * n2, cudaMemcpyHostToDevice, h->stream));
      d_add2 = h->st_gues2.p; d_an2 = h->st_anal2.p;
    }
  }
  DevBuf<int> &d_sh = h->so_tmp;   // (scratch of set_obs, free between calls)
  const int *d_ishuf = nullptr, *d_inv = nullptr;
  std::vector<int> inv;             // destination member of source member ms (register-resident kernel); empty: not a permutation
  if (ishuf) {
    inv.assign((size_t)k, -1);
    bool perm = true;
    for (int m = 0; m < k; ++m) {
      if (inv[ishuf[m] - 1] >= 0) perm = false;
      inv[ishuf[m] - 1] = m;
    }
    if (!perm) inv.clear();
    CK(d_sh.ensure(2 * (size_t)k));
    CK(cudaMemcpyAsync(d_sh.p, ishuf, sizeof(int) * k, cudaMemcpyHostToDevice, h->stream));
    if (!inv.empty()) CK(cudaMemcpyAsync(d_sh.p + k, inv.data(), sizeof(int) * k, cudaMemcpyHostToDevice, h->stream));
    d_ishuf = d_sh.p;
    d_inv = inv.empty() ? nullptr : d_sh.p + k;
  }
  // register-resident variant (one read of the additive members): opt-in, LETKF_B200_ADDINFL_REG=1 -- at 158 registers its
  // occupancy is too low to beat the two-pass kernel on B200 (measured), kept for the A/B
  static const bool reg_opt = std::getenv("LETKF_B200_ADDINFL_REG") != nullptr;
  const bool reg_path = k <= 64 && reg_opt && (!ishuf || d_inv);
  const int q_lo = c.iv3d_q - 1, q_hi = c.iv3d_qg - 1;   // (the moisture variables q, qc, qr, qi, qs, qg are contiguous)
  // background mean = slot MEMBER of gues3d on the device, or the nv3d planes staged above
  const double *d_gm = !q_ratio ? nullptr : host ? h->st_rtps.p : gues3d;
  const size_t gm_vs = host ? sl : sl * nens, gm_off = host ? 0 : (size_t)k * sl;
  if (reg_path && k <= 32)
    additive_inflation_reg_kernel<32><<<(unsigned)((o3 + 127) / 128), 128, 0, h->stream>>>(k, nens, nij, sl, c.nv3d, d_add3, d_an3, d_gm, gm_vs,
                                                                                          gm_off, d_w, d_inv, infl_add, q_lo, q_hi);
  else if (reg_path)
    additive_inflation_reg_kernel<64><<<(unsigned)((o3 + 127) / 128), 128, 0, h->stream>>>(k, nens, nij, sl, c.nv3d, d_add3, d_an3, d_gm, gm_vs,
                                                                                          gm_off, d_w, d_inv, infl_add, q_lo, q_hi);
  else
    additive_inflation_kernel<<<(unsigned)((o3 + 255) / 256), 256, 0, h->stream>>>(k, nens, nij, sl, c.nv3d, d_add3, d_an3, d_gm, gm_vs, gm_off, d_w,
                                                                                   d_ishuf, infl_add, q_lo, q_hi);
  if (two)
    additive_inflation_kernel<<<(unsigned)(((size_t)nij * c.nv2d + 255) / 256), 256, 0, h->stream>>>(k, nens, nij, (size_t)nij, c.nv2d, d_add2, d_an2,
                                                                                                     nullptr, 0, 0, d_w, d_ishuf, infl_add, -1, -2);
  CK(cudaGetLastError());
  if (host) {
    CK(cudaMemcpyAsync(anal3d, d_an3, sizeof(double) * n3, cudaMemcpyDeviceToHost, h->stream));
    if (two) CK(cudaMemcpyAsync(anal2d, d_an2, sizeof(double) * n2, cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  return LETKF_B200_OK;
}

void letkf_b200_thermo_defaults(letkf_b200_thermo *t) {   // SCALE-RM scale_const / scale_tracer values
  std::memset(t, 0, sizeof(*t));
  t->Rdry = 287.04;
  t->Rvap = 461.46;
  t->CVdry = 1004.64 - 287.04;
  t->PRE00 = 1.0e5;
  t->TRACER_CV[0] = 1845.60 - 461.46;   // vapour
  t->TRACER_CV[1] = t->TRACER_CV[2] = 4218.0;                    // cloud water, rain
  t->TRACER_CV[3] = t->TRACER_CV[4] = t->TRACER_CV[5] = 2006.0;  // ice, snow, graupel
  t->POSITIVE_DEFINITE_Q = 0;
  t->POSITIVE_DEFINITE_QHYD = 0;
}

int letkf_b200_state_trans(letkf_b200_handle *h, const letkf_b200_thermo *t, int inverse, double *v3dg, int mem_space) {
  if (!h || !t || !v3dg) return LETKF_B200_EINVAL;
  const letkf_b200_config &c = h->cfg;
  if (c.nv3d < 6 || c.iv3d_q != 6 || c.iv3d_p != 5) return fail(h, LETKF_B200_EINVAL, "state_trans needs the u,v,w,T,p,q.. variable order");
  CK(cudaSetDevice(h->device));
  const long long npts = (long long)c.nlev * c.nlon * c.nlat;
  const size_t n = (size_t)npts * c.nv3d;
  const bool host = mem_space != LETKF_B200_MEM_DEVICE;
  StateTransParams P;
  std::memset(&P, 0, sizeof(P));
  P.Rdry = t->Rdry; P.Rvap = t->Rvap; P.CVdry = t->CVdry; P.PRE00 = t->PRE00;
  for (int i = 0; i < 16; ++i) P.tracer_cv[i] = t->TRACER_CV[i];
  P.pos_q = t->POSITIVE_DEFINITE_Q; P.pos_qhyd = t->POSITIVE_DEFINITE_QHYD;
  P.nv3d = c.nv3d; P.iv3d_q = c.iv3d_q; P.npts = npts;
  if (host) {
    CK(h->cb[0].ensure(n));
    CK(cudaMemcpyAsync(h->cb[0].p, v3dg, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
    P.v = h->cb[0].p;
  } else {
    P.v = v3dg;
  }
  state_trans_kernel<<<(unsigned)((npts + 255) / 256), 256, 0, h->stream>>>(P, inverse ? 1 : 0);
  CK(cudaGetLastError());
  if (host) {
    CK(cudaMemcpyAsync(v3dg, P.v, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return LETKF_B200_OK;
}

int letkf_b200_nij1(const letkf_b200_handle *h, int np, int myrank_e, int32_t *nij1, int32_t *nij1max) {
  if (!h || np < 1 || myrank_e < 0 || myrank_e >= np) return LETKF_B200_EINVAL;
  const int tot = h->cfg.nlon * h->cfg.nlat;
  const int i = tot % np;   // common_mpi_scale.f90:264-271
  const int mx = (tot - i) / np + 1;
  if (nij1max) *nij1max = mx;
  if (nij1) *nij1 = myrank_e < i ? mx : mx - 1;
  return LETKF_B200_OK;
}

static TransposeDims make_dims(const letkf_b200_handle *h, int np) {
  TransposeDims d;
  const letkf_b200_config &c = h->cfg;
  d.nlon = c.nlon; d.nlat = c.nlat; d.nlev = c.nlev; d.nv3d = c.nv3d; d.nv2d = c.nv2d; d.np = np;
  const int tot = c.nlon * c.nlat, i = tot % np;
  d.nij1max = (tot - i) / np + 1;
  d.nlevall = c.nlev * c.nv3d + c.nv2d;
  return d;
}

// tiled pack (dir 0) / unpack (dir 1) with the optional fused state transform
static int grd_buf_tiled(letkf_b200_handle *h, int np, const letkf_b200_thermo *t, int dir, double *v3dg, double *v2dg,
                         double *buf) {
  CK(cudaSetDevice(h->device));
  const letkf_b200_config &c = h->cfg;
  if (t && (c.nv3d < 6 || c.iv3d_q != 6 || c.iv3d_p != 5))
    return fail(h, LETKF_B200_EINVAL, "state_trans needs the u,v,w,T,p,q.. variable order");
  const TransposeDims d = make_dims(h, np);
  StateTransParams P;
  std::memset(&P, 0, sizeof(P));
  if (t) {
    P.Rdry = t->Rdry; P.Rvap = t->Rvap; P.CVdry = t->CVdry; P.PRE00 = t->PRE00;
    for (int i = 0; i < 16; ++i) P.tracer_cv[i] = t->TRACER_CV[i];
    P.pos_q = t->POSITIVE_DEFINITE_Q; P.pos_qhyd = t->POSITIVE_DEFINITE_QHYD;
  }
  P.nv3d = c.nv3d; P.iv3d_q = c.iv3d_q;
  const size_t smem = (size_t)c.nv3d * 32 * 33 * sizeof(double);
  // (the attribute is per device: set it on every call, a process may hold handles on several GPUs)
  CK(cudaFuncSetAttribute(grd_buf_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 32 * 33 * (int)sizeof(double)));
  const dim3 grid((unsigned)((d.nij1max + 31) / 32), (unsigned)((c.nlev + 31) / 32), (unsigned)np);
  grd_buf_tiled_kernel<<<grid, dim3(32, 8), smem, h->stream>>>(d, P, t ? 1 : 0, dir, v3dg, buf);
  if (c.nv2d > 0 && v2dg) grd_buf_2d_kernel<<<h->num_sms * 4, 256, 0, h->stream>>>(d, dir, v2dg, buf);
  CK(cudaGetLastError());
  return LETKF_B200_OK;
}

int letkf_b200_grd_to_buf(letkf_b200_handle *h, int np, const double *v3dg, const double *v2dg, double *bufs) {
  if (!h || np < 1 || !v3dg || !bufs) return LETKF_B200_EINVAL;
  return grd_buf_tiled(h, np, nullptr, 0, const_cast<double *>(v3dg), const_cast<double *>(v2dg), bufs);
}
int letkf_b200_buf_to_grd(letkf_b200_handle *h, int np, const double *bufr, double *v3dg, double *v2dg) {
  if (!h || np < 1 || !v3dg || !bufr) return LETKF_B200_EINVAL;
  return grd_buf_tiled(h, np, nullptr, 1, v3dg, v2dg, const_cast<double *>(bufr));
}
int letkf_b200_grd_to_buf_trans(letkf_b200_handle *h, int np, const letkf_b200_thermo *t, const double *v3dg,
                                const double *v2dg, double *bufs) {
  if (!h || np < 1 || !v3dg || !bufs) return LETKF_B200_EINVAL;
  return grd_buf_tiled(h, np, t, 0, const_cast<double *>(v3dg), const_cast<double *>(v2dg), bufs);
}
int letkf_b200_buf_to_grd_trans(letkf_b200_handle *h, int np, const letkf_b200_thermo *t, const double *bufr, double *v3dg,
                                double *v2dg) {
  if (!h || np < 1 || !v3dg || !bufr) return LETKF_B200_EINVAL;
  return grd_buf_tiled(h, np, t, 1, v3dg, v2dg, const_cast<double *>(bufr));
}
int letkf_b200_buf_to_ens(letkf_b200_handle *h, int np, int myrank_e, int nens, int mstart, int mend,
                          const double *bufr, double *v3d, double *v2d) {
  if (!h || np < 1 || !v3d || !bufr || mstart < 1 || mend < mstart || mend > nens) return LETKF_B200_EINVAL;
  CK(cudaSetDevice(h->device));
  const TransposeDims d = make_dims(h, np);
  int32_t nij1;
  letkf_b200_nij1(h, np, myrank_e, &nij1, nullptr);
  buf_to_ens_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(d, nij1, nens, mstart, mend - mstart + 1, bufr, v3d, v2d, 0);
  CK(cudaGetLastError());
  return LETKF_B200_OK;
}
int letkf_b200_ens_to_buf(letkf_b200_handle *h, int np, int myrank_e, int nens, int mstart, int mend,
                          const double *v3d, const double *v2d, double *bufs) {
  if (!h || np < 1 || !v3d || !bufs || mstart < 1 || mend < mstart || mend > nens) return LETKF_B200_EINVAL;
  CK(cudaSetDevice(h->device));
  const TransposeDims d = make_dims(h, np);
  int32_t nij1;
  letkf_b200_nij1(h, np, myrank_e, &nij1, nullptr);
  buf_to_ens_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(d, nij1, nens, mstart, mend - mstart + 1, bufs,
                                                          const_cast<double *>(v3d), const_cast<double *>(v2d), 1);
  CK(cudaGetLastError());
  return LETKF_B200_OK;
}

// ---- one-pass transposes over peer memory ------------------------------------------------------------------
int letkf_b200_peer_export(letkf_b200_handle *h, const void *devptr, letkf_b200_ipc *out) {
  if (!h || !devptr || !out) return LETKF_B200_EINVAL;
  CK(cudaSetDevice(h->device));
  CUdeviceptr base = 0;
  size_t size = 0;
  {   // the IPC handle names the whole allocation: find its base (driver entry point through the runtime, no -lcuda)
    typedef CUresult (*range_fn)(CUdeviceptr *, size_t *, CUdeviceptr);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr));
    if (!fn || ((range_fn)fn)(&base, &size, (CUdeviceptr)devptr) != CUDA_SUCCESS)
      return fail(h, LETKF_B200_ECUDA, "peer_export: cuMemGetAddressRange failed (not a device allocation?)");
  }
  cudaIpcMemHandle_t hd;
  CK(cudaIpcGetMemHandle(&hd, (void *)base));
  static_assert(sizeof(hd) == 64, "cudaIpcMemHandle_t is 64 bytes");
  std::memcpy(out->handle, &hd, 64);
  out->offset = (uint64_t)((CUdeviceptr)devptr - base);
  return LETKF_B200_OK;
}
int letkf_b200_peer_open(letkf_b200_handle *h, const letkf_b200_ipc *in, void **mapped) {
  if (!h || !in || !mapped) return LETKF_B200_EINVAL;
  CK(cudaSetDevice(h->device));
  const std::string key((const char *)in->handle, 64);
  auto it = h->ipc_open.find(key);
  void *base = nullptr;
  if (it != h->ipc_open.end()) {
    base = it->second;
  } else {
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, in->handle, 64);
    CK(cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess));
    h->ipc_open[key] = base;
  }
  *mapped = (char *)base + in->offset;
  return LETKF_B200_OK;
}

static int grd_ens_p2p(letkf_b200_handle *h, int np, int myrank, int nens, int dir, int slot0, int npeers,
                       const letkf_b200_thermo *t, const double *src3, const double *src2, double *const *peer3,
                       double *const *peer2) {
  CK(cudaSetDevice(h->device));
  const letkf_b200_config &c = h->cfg;
  if (np < 1 || np > 16 || npeers < 1 || npeers > np) return fail(h, LETKF_B200_EINVAL, "p2p transposes support 1..16 ranks");
  if (t && (c.nv3d < 6 || c.iv3d_q != 6 || c.iv3d_p != 5))
    return fail(h, LETKF_B200_EINVAL, "state_trans needs the u,v,w,T,p,q.. variable order");
  const TransposeDims d = make_dims(h, np);
  StateTransParams P;
  std::memset(&P, 0, sizeof(P));
  if (t) {
    P.Rdry = t->Rdry; P.Rvap = t->Rvap; P.CVdry = t->CVdry; P.PRE00 = t->PRE00;
    for (int i = 0; i < 16; ++i) P.tracer_cv[i] = t->TRACER_CV[i];
    P.pos_q = t->POSITIVE_DEFINITE_Q; P.pos_qhyd = t->POSITIVE_DEFINITE_QHYD;
  }
  P.nv3d = c.nv3d; P.iv3d_q = c.iv3d_q;
  PeerPtrs pp;
  std::memset(&pp, 0, sizeof(pp));
  for (int i = 0; i < npeers; ++i) {
    if (!peer3[i]) return fail(h, LETKF_B200_EINVAL, "p2p transpose: null peer pointer");
    pp.p3[i] = peer3[i];
    pp.p2[i] = (peer2 && c.nv2d > 0) ? peer2[i] : nullptr;
  }
  static const int variant = [] { const char *e = std::getenv("LETKF_B200_P2P_TILE"); return e ? std::atoi(e) : 16; }();
  if (variant == 32) {   // A/B aid: the register-staged 32-level tile of the first version
    const size_t smem = (size_t)c.nv3d * 32 * 33 * sizeof(double);
    CK(cudaFuncSetAttribute(grd_ens_p2p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 32 * 33 * (int)sizeof(double)));
    const dim3 grid((unsigned)((d.nij1max + 31) / 32), (unsigned)((c.nlev + 31) / 32), (unsigned)npeers);
    grd_ens_p2p_kernel<<<grid, dim3(32, 8), smem, h->stream>>>(d, P, t ? 1 : 0, dir, myrank, nens, slot0, src3, pp);
  } else {
    constexpr int KT = 16;
    if ((d.nij1max + 31) / 32 > 65535) return fail(h, LETKF_B200_EINVAL, "p2p transposes: more than 2 M columns per rank");
    const size_t smem = (size_t)c.nv3d * 32 * (KT + 1) * sizeof(double);
    CK(cudaFuncSetAttribute(grd_ens_p2p_async_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 32 * (KT + 1) * (int)sizeof(double)));
    const dim3 grid((unsigned)((c.nlev + KT - 1) / KT), (unsigned)((d.nij1max + 31) / 32), (unsigned)npeers);
    grd_ens_p2p_async_kernel<KT><<<grid, 256, smem, h->stream>>>(d, P, t ? 1 : 0, dir, myrank, nens, slot0, src3, pp);
  }
  if (c.nv2d > 0 && src2 && peer2)
    grd_ens_p2p_2d_kernel<<<h->num_sms * 4, 256, 0, h->stream>>>(d, dir, myrank, nens, slot0, npeers, src2, pp);
  CK(cudaGetLastError());
  return LETKF_B200_OK;
}

int letkf_b200_scatter_grd_p2p(letkf_b200_handle *h, int np, int myrank_e, int nens, int mslot, const letkf_b200_thermo *t,
                               const double *v3dg, const double *v2dg, double *const *peer_v3d, double *const *peer_v2d) {
  if (!h || !v3dg || !peer_v3d || mslot < 1 || mslot > nens) return LETKF_B200_EINVAL;
  return grd_ens_p2p(h, np, myrank_e, nens, 0, mslot - 1, np, t, v3dg, v2dg, peer_v3d, peer_v2d);
}
int letkf_b200_gather_grd_p2p(letkf_b200_handle *h, int np, int myrank_e, int nens, int mstart, int mend,
                              const letkf_b200_thermo *t, const double *v3d, const double *v2d, double *const *peer_v3dg,
                              double *const *peer_v2dg) {
  if (!h || !v3d || !peer_v3dg || mstart < 1 || mend < mstart || mend > nens || mend - mstart + 1 > np) return LETKF_B200_EINVAL;
  return grd_ens_p2p(h, np, myrank_e, nens, 1, mstart - 1, mend - mstart + 1, t, v3d, v2d, peer_v3dg, peer_v2dg);
}

// ---- radar observation operator ------------------------------------------------------------------------------
void letkf_b200_radar_config_defaults(letkf_b200_radar_config *r) {   // common_nml.f90:257-272
  std::memset(r, 0, sizeof(*r));
  r->METHOD_REF_CALC = 3;
  r->USE_TERMINAL_VELOCITY = 0;
  r->MIN_RADAR_REF_DBZ = 0.0;
  r->LOW_REF_SHIFT = 0.0;
  r->RADAR_ZMAX = 99.0e3;
  r->nv3dd = 13;
  r->KHALO = 2;
}

int letkf_b200_obsope_radar(letkf_b200_handle *h, const letkf_b200_radar_config *r, int nobs, const int32_t *elm,
                            const double *ril, const double *rjl, const double *lon, const double *lat, const double *lev,
                            const double *rotc, int nmem, const double *const *v3dgh, int ld_out, double *yobs, int32_t *qc,
                            int mem_space) {
  if (!h || !r || nobs < 0 || nmem < 1 || ld_out < nmem || !v3dgh) return LETKF_B200_EINVAL;
  if (nobs == 0) return LETKF_B200_OK;
  if (!elm || !ril || !rjl || !lon || !lat || !lev || !yobs || !qc) return LETKF_B200_EINVAL;
  if (r->METHOD_REF_CALC < 1 || r->METHOD_REF_CALC > 3)
    return fail(h, LETKF_B200_EINVAL, "Not recognized method for radar reflectivity and wind computation");
  if (r->nv3dd < 13 || r->nlev + 2 * r->KHALO > r->nlevh) return fail(h, LETKF_B200_EINVAL, "obsope_radar: inconsistent grid sizes");
  CK(cudaSetDevice(h->device));
  const bool host = mem_space != LETKF_B200_MEM_DEVICE;
  RadarParams P;
  std::memset(&P, 0, sizeof(P));
  P.c.method = r->METHOD_REF_CALC; P.c.use_tv = r->USE_TERMINAL_VELOCITY;
  P.c.min_ref = std::pow(10.0, r->MIN_RADAR_REF_DBZ / 10.0);   // common_obs_scale.f90:251 (host libm, like the reference)
  P.c.min_ref_dbz = r->MIN_RADAR_REF_DBZ; P.c.low_ref_shift = r->LOW_REF_SHIFT;
  P.nobs = nobs; P.nmem = nmem; P.nlevh = r->nlevh; P.nlonh = r->nlonh; P.nlath = r->nlath; P.nlev = r->nlev; P.khalo = r->KHALO;
  P.nv3dd = r->nv3dd; P.ld_out = ld_out; P.zmax = r->RADAR_ZMAX;
  P.radar_lon = r->radar_lon; P.radar_lat = r->radar_lat; P.radar_z = r->radar_z;
  const size_t gsz = (size_t)r->nlevh * r->nlonh * r->nlath * r->nv3dd, no = (size_t)nobs * ld_out;
  DevBuf<double> d_geo, d_grid, d_y;
  DevBuf<int> d_elm, d_qc;
  DevBuf<const double *> d_ptr;
  CK(d_ptr.ensure(nmem));
  std::vector<const double *> ptrs(nmem);
  if (host) {
    CK(d_geo.ensure((size_t)nobs * 7)); CK(d_elm.ensure(nobs)); CK(d_y.ensure(no)); CK(d_qc.ensure(no));
    CK(d_grid.ensure(gsz * nmem));
    const double *src[5] = {ril, rjl, lon, lat, lev};
    for (int i = 0; i < 5; ++i)
      CK(cudaMemcpyAsync(d_geo.p + (size_t)i * nobs, src[i], sizeof(double) * nobs, cudaMemcpyHostToDevice, h->stream));
    if (rotc) CK(cudaMemcpyAsync(d_geo.p + (size_t)5 * nobs, rotc, sizeof(double) * 2 * nobs, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_elm.p, elm, sizeof(int) * nobs, cudaMemcpyHostToDevice, h->stream));
    for (int m = 0; m < nmem; ++m) {
      CK(cudaMemcpyAsync(d_grid.p + gsz * m, v3dgh[m], sizeof(double) * gsz, cudaMemcpyHostToDevice, h->stream));
      ptrs[m] = d_grid.p + gsz * m;
    }
    P.ril = d_geo.p; P.rjl = d_geo.p + nobs; P.lon = d_geo.p + 2 * (size_t)nobs; P.lat = d_geo.p + 3 * (size_t)nobs;
    P.lev = d_geo.p + 4 * (size_t)nobs; P.rotc = rotc ? d_geo.p + 5 * (size_t)nobs : nullptr;
    P.elm = d_elm.p; P.yobs = d_y.p; P.qc = d_qc.p;
  } else {
    for (int m = 0; m < nmem; ++m) ptrs[m] = v3dgh[m];
    P.ril = ril; P.rjl = rjl; P.lon = lon; P.lat = lat; P.lev = lev; P.rotc = rotc; P.elm = elm; P.yobs = yobs; P.qc = qc;
  }
  CK(cudaMemcpyAsync(d_ptr.p, ptrs.data(), sizeof(double *) * nmem, cudaMemcpyHostToDevice, h->stream));
  P.v3dgh = d_ptr.p;
  obsope_radar_kernel<<<dim3((unsigned)((nobs + 127) / 128), (unsigned)nmem), 128, 0, h->stream>>>(P);
  CK(cudaGetLastError());
  if (host) {
    CK(cudaMemcpyAsync(yobs, d_y.p, sizeof(double) * no, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(qc, d_qc.p, sizeof(int) * no, cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  d_geo.release(); d_grid.release(); d_y.release(); d_elm.release(); d_qc.release(); d_ptr.release();
  return LETKF_B200_OK;
}

// ---- conventional observation operator and monit_obs ---------------------------------------------------------------
void letkf_b200_conv_config_defaults(letkf_b200_conv_config *c) {
  std::memset(c, 0, sizeof(*c));
  c->KHALO = 2;
  c->nv3dd = 13;
  c->nv2dd = 7;
  c->stggrd = 0;
  c->PS_ADJUST_THRES = 100.0;   // common_nml.f90:148
}

int letkf_b200_obsope_conv(letkf_b200_handle *h, const letkf_b200_conv_config *r, int nobs, const int32_t *elm, const double *ril,
                           const double *rjl, const double *lev, const double *rotc, int nmem, const double *const *v3dgh,
                           const double *const *v2dgh, int ld_out, double *yobs, int32_t *qc, int mem_space) {
  if (!h || !r || nobs < 0 || nmem < 1 || ld_out < nmem || !v3dgh || !v2dgh) return LETKF_B200_EINVAL;
  if (nobs == 0) return LETKF_B200_OK;
  if (!elm || !ril || !rjl || !lev || !yobs || !qc) return LETKF_B200_EINVAL;
  if (r->nv3dd < 13 || r->nv2dd < 7 || r->nlev + 2 * r->KHALO > r->nlevh)
    return fail(h, LETKF_B200_EINVAL, "obsope_conv: inconsistent grid sizes");
  CK(cudaSetDevice(h->device));
  const bool host = mem_space != LETKF_B200_MEM_DEVICE;
  ConvParams P;
  std::memset(&P, 0, sizeof(P));
  P.nobs = nobs; P.nmem = nmem; P.nlevh = r->nlevh; P.nlonh = r->nlonh; P.nlath = r->nlath; P.nlev = r->nlev; P.khalo = r->KHALO;
  P.ld_out = ld_out; P.stggrd = r->stggrd; P.ps_thres = r->PS_ADJUST_THRES;
  const size_t g3 = (size_t)r->nlevh * r->nlonh * r->nlath * r->nv3dd, g2 = (size_t)r->nlonh * r->nlath * r->nv2dd;
  const size_t no = (size_t)nobs * ld_out;
  DevBuf<double> d_geo, d_grid, d_y;
  DevBuf<int> d_elm, d_qc;
  DevBuf<const double *> d_ptr;
  CK(d_ptr.ensure(2 * (size_t)nmem));
  std::vector<const double *> ptrs(2 * (size_t)nmem);
  if (host) {
    CK(d_geo.ensure((size_t)nobs * 5)); CK(d_elm.ensure(nobs)); CK(d_y.ensure(no)); CK(d_qc.ensure(no));
    CK(d_grid.ensure((g3 + g2) * nmem));
    const double *src[3] = {ril, rjl, lev};
    for (int i = 0; i < 3; ++i)
      CK(cudaMemcpyAsync(d_geo.p + (size_t)i * nobs, src[i], sizeof(double) * nobs, cudaMemcpyHostToDevice, h->stream));
    if (rotc) CK(cudaMemcpyAsync(d_geo.p + (size_t)3 * nobs, rotc, sizeof(double) * 2 * nobs, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_elm.p, elm, sizeof(int) * nobs, cudaMemcpyHostToDevice, h->stream));
    for (int m = 0; m < nmem; ++m) {
      double *b3 = d_grid.p + (g3 + g2) * m, *b2 = b3 + g3;
      CK(cudaMemcpyAsync(b3, v3dgh[m], sizeof(double) * g3, cudaMemcpyHostToDevice, h->stream));
      CK(cudaMemcpyAsync(b2, v2dgh[m], sizeof(double) * g2, cudaMemcpyHostToDevice, h->stream));
      ptrs[m] = b3;
      ptrs[nmem + m] = b2;
    }
    P.ril = d_geo.p; P.rjl = d_geo.p + nobs; P.lev = d_geo.p + 2 * (size_t)nobs; P.rotc = rotc ? d_geo.p + 3 * (size_t)nobs : nullptr;
    P.elm = d_elm.p; P.yobs = d_y.p; P.qc = d_qc.p;
  } else {
    for (int m = 0; m < nmem; ++m) {
      ptrs[m] = v3dgh[m];
      ptrs[nmem + m] = v2dgh[m];
    }
    P.ril = ril; P.rjl = rjl; P.lev = lev; P.rotc = rotc; P.elm = elm; P.yobs = yobs; P.qc = qc;
  }
  CK(cudaMemcpyAsync(d_ptr.p, ptrs.data(), sizeof(double *) * 2 * nmem, cudaMemcpyHostToDevice, h->stream));
  P.v3dgh = d_ptr.p;
  P.v2dgh = d_ptr.p + nmem;
  obsope_conv_kernel<<<dim3((unsigned)((nobs + 127) / 128), (unsigned)nmem), 128, 0, h->stream>>>(P);
  CK(cudaGetLastError());
  if (host) {
    CK(cudaMemcpyAsync(yobs, d_y.p, sizeof(double) * no, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(qc, d_qc.p, sizeof(int) * no, cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  d_geo.release(); d_grid.release(); d_y.release(); d_elm.release(); d_qc.release(); d_ptr.release();
  return LETKF_B200_OK;
}

int letkf_b200_monit_obs_set(letkf_b200_handle *h, const letkf_b200_conv_config *conv, const letkf_b200_radar_config *radar, int nobs,
                             const int32_t *elm, const double *ril, const double *rjl, const double *lon, const double *lat,
                             const double *lev, const double *dat, const double *dif, const double *rotc, double t_range,
                             const double *v3dgh, const double *v2dgh, double *ohx, int32_t *oqc, int mem_space) {
  if (!h || (!conv) == (!radar) || nobs < 0) return LETKF_B200_EINVAL;   // exactly one format
  if (nobs == 0) return LETKF_B200_OK;
  if (!elm || !ril || !rjl || !lev || !dat || !v3dgh || !ohx || !oqc || (conv && !v2dgh) || (radar && (!lon || !lat))) return LETKF_B200_EINVAL;
  CK(cudaSetDevice(h->device));
  const bool host = mem_space != LETKF_B200_MEM_DEVICE;
  const int nlevh = conv ? conv->nlevh : radar->nlevh, nlonh = conv ? conv->nlonh : radar->nlonh, nlath = conv ? conv->nlath : radar->nlath;
  const size_t g3 = (size_t)nlevh * nlonh * nlath * (conv ? conv->nv3dd : radar->nv3dd);
  const size_t g2 = conv ? (size_t)nlonh * nlath * conv->nv2dd : 0;
  // everything on the device: inputs are staged when they are host arrays, then the operators run in device mode
  DevBuf<double> d_in, d_grid, d_y, d_ohx;
  DevBuf<int> d_elm, d_q, d_oqc;
  const double *p_ril = ril, *p_rjl = rjl, *p_lon = lon, *p_lat = lat, *p_lev = lev, *p_dat = dat, *p_dif = dif, *p_rotc = rotc;
  const double *p_g3 = v3dgh, *p_g2 = v2dgh;
  const int *p_elm = elm;
  double *p_ohx = ohx;
  int *p_oqc = oqc;
  if (host) {
    CK(d_in.ensure((size_t)nobs * 9)); CK(d_elm.ensure(nobs)); CK(d_grid.ensure(g3 + g2)); CK(d_ohx.ensure(nobs)); CK(d_oqc.ensure(nobs));
    const double *src[7] = {ril, rjl, lon, lat, lev, dat, dif};
    const double **dstp[7] = {&p_ril, &p_rjl, &p_lon, &p_lat, &p_lev, &p_dat, &p_dif};
    for (int i = 0; i < 7; ++i) {
      if (!src[i]) continue;
      CK(cudaMemcpyAsync(d_in.p + (size_t)i * nobs, src[i], sizeof(double) * nobs, cudaMemcpyHostToDevice, h->stream));
      *dstp[i] = d_in.p + (size_t)i * nobs;
    }
    if (rotc) {
      CK(cudaMemcpyAsync(d_in.p + (size_t)7 * nobs, rotc, sizeof(double) * 2 * nobs, cudaMemcpyHostToDevice, h->stream));
      p_rotc = d_in.p + (size_t)7 * nobs;
    }
    CK(cudaMemcpyAsync(d_elm.p, elm, sizeof(int) * nobs, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_grid.p, v3dgh, sizeof(double) * g3, cudaMemcpyHostToDevice, h->stream));
    if (g2) CK(cudaMemcpyAsync(d_grid.p + g3, v2dgh, sizeof(double) * g2, cudaMemcpyHostToDevice, h->stream));
    p_elm = d_elm.p; p_g3 = d_grid.p; p_g2 = d_grid.p + g3; p_ohx = d_ohx.p; p_oqc = d_oqc.p;
    CK(cudaStreamSynchronize(h->stream));
  }
  CK(d_y.ensure(nobs)); CK(d_q.ensure(nobs));
  int rc;
  if (conv) {
    rc = letkf_b200_obsope_conv(h, conv, nobs, p_elm, p_ril, p_rjl, p_lev, p_rotc, 1, &p_g3, &p_g2, 1, d_y.p, d_q.p, LETKF_B200_MEM_DEVICE);
  } else {
    letkf_b200_radar_config rr = *radar;
    rr.RADAR_ZMAX = 1.0e300;   // monit_obs has no RADAR_ZMAX test (common_obs_scale.f90:1545-1555; obsope_cal has: obsope_tools.f90:476)
    rc = letkf_b200_obsope_radar(h, &rr, nobs, p_elm, p_ril, p_rjl, p_lon, p_lat, p_lev, p_rotc, 1, &p_g3, 1, d_y.p, d_q.p, LETKF_B200_MEM_DEVICE);
  }
  if (rc != LETKF_B200_OK) return rc;
  monit_ohx_kernel<<<(nobs + 255) / 256, 256, 0, h->stream>>>(nobs, p_dat, p_dif, t_range, d_y.p, d_q.p, p_ohx, p_oqc);
  CK(cudaGetLastError());
  if (host) {
    CK(cudaMemcpyAsync(ohx, p_ohx, sizeof(double) * nobs, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(oqc, p_oqc, sizeof(int) * nobs, cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  return LETKF_B200_OK;
}

}  // extern "C"
