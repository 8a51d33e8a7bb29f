// oracle_das.cpp -- CPU restatement of das_letkf, obs_local and the bucket tables.
//
// TEST INFRASTRUCTURE ONLY (see letkf_oracle.h).  PARITY UNPINNED by the reference.
//
// Single sorting mesh over the whole horizontal plane (the PRC_NUM_X = PRC_NUM_Y = 1 view
// of the reference): rank_i = rank_j = 0 in rij_g2l (common_scale.f90:1683-1697), the
// extended mesh carries ngrdsch empty halo buckets on every side.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "letkf_oracle.h"

namespace {

const int NOBTYPE = LETKF_B200_NOBTYPE;
const int NID_OBS = LETKF_B200_NID_OBS;
const int N_SEARCH_INCR = 8;   // letkf_tools.f90:44

// raw element ids, common_obs_scale.f90:45-77
const int ID_PS = 14593, ID_RAIN = 19999, ID_REF = 4001, ID_RE0 = 4004, ID_VR = 4002;
const int ELEM_UID[NID_OBS] = {2819, 2820, 3073, 3074, 3330, 3331, 14593, 19999,
                               4001, 4004, 4002, 4003, 8800, 99991, 99992, 99993};

int uid_obs(int elm) {   // common_obs_scale.f90:171-211
  for (int i = 0; i < NID_OBS; ++i)
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

struct Mesh {   // obs_grid_type, letkf_obs.f90:47-65 (single-subdomain fields only)
  int ngrd_i, ngrd_j, ngrdsch_i, ngrdsch_j, ngrdext_i, ngrdext_j, tot_ext;
  double grdspc_i, grdspc_j;
  std::vector<int> n_ext;    // (ngrdext_i, ngrdext_j)
  std::vector<int> ac_ext;   // (0:ngrdext_i, ngrdext_j)
  int &N(int i, int j) { return n_ext[(i - 1) + (size_t)(j - 1) * ngrdext_i]; }
  int AC(int i, int j) const { return ac_ext[i + (size_t)(j - 1) * (ngrdext_i + 1)]; }
  int &ACw(int i, int j) { return ac_ext[i + (size_t)(j - 1) * (ngrdext_i + 1)]; }
};

}  // namespace

struct oracle_state {
  letkf_b200_config cfg;
  int quirk_ij = 0;
  // ctype tables, letkf_obs.f90:33-41
  int nctype = 0;
  int ctype_elmtyp[NID_OBS][NOBTYPE];
  std::vector<int> elm_ctype, elm_u_ctype, typ_ctype;
  std::vector<double> hori_loc_ctype, vert_loc_ctype;
  std::vector<Mesh> obsgrd;
  int nobstotal = 0;
  // obsda_sort + obs(:) fields in sorted order
  int nensobs = 0;
  std::vector<int> sorted_to_orig, s_elm, s_typ;
  std::vector<double> s_ri, s_rj, s_lev, s_dat, s_err, s_val, s_ensval;
  // merge tables, letkf_tools.f90:167-192
  std::vector<int> n_merge;
  std::vector<std::vector<int>> ic_merge;
  int n_merge_max = 1;
  bool radar_only = true;
  // grid
  int nij1 = 0;
  std::vector<double> rig1, rjg1, hgt1;
  // variable localisation, letkf_tools.f90:130-163
  double var_local[LETKF_B200_MAX_NV][LETKF_B200_NID_VARLOCAL];
  int var_local_n2nc[LETKF_B200_MAX_NV], var_local_n2n[LETKF_B200_MAX_NV], n2nc_max = 1;
};

namespace {

// ij_obsgrd_ext, letkf_obs.f90:1209-1227 (rank 0: ril = ri).  0.5 is a default-REAL
// literal there; it is exact in both precisions.
void ij_obsgrd_ext(const oracle_state &s, int ic, double ri, double rj, int &ogi, int &ogj) {
  const Mesh &g = s.obsgrd[ic];
  ogi = (int)std::ceil((ri - (double)s.cfg.IHALO - 0.5) * (double)g.ngrd_i / (double)s.cfg.nlon) +
        g.ngrdsch_i;
  ogj = (int)std::ceil((rj - (double)s.cfg.JHALO - 0.5) * (double)g.ngrd_j / (double)s.cfg.nlat) +
        g.ngrdsch_j;
}
// ij_obsgrd, letkf_obs.f90:1187-1204 (used for binning; quirk: ngrd_i in the j formula)
void ij_obsgrd(const oracle_state &s, int ic, double ri, double rj, int &ogi, int &ogj) {
  const Mesh &g = s.obsgrd[ic];
  ogi = (int)std::ceil((ri - (double)s.cfg.IHALO - 0.5) * (double)g.ngrd_i / (double)s.cfg.nlon);
  const int nj = s.quirk_ij ? g.ngrd_i : g.ngrd_j;
  ogj = (int)std::ceil((rj - (double)s.cfg.JHALO - 0.5) * (double)nj / (double)s.cfg.nlat);
}
// obs_choose_ext, letkf_obs.f90:1262-1285; appends 0-based sorted indices
void obs_choose_ext(const oracle_state &s, int ic, int imin, int imax, int jmin, int jmax, int &nn,
                    std::vector<int> *nobs_use) {
  const Mesh &g = s.obsgrd[ic];
  if (imin > imax || jmin > jmax) return;
  if (g.tot_ext == 0) return;
  for (int j = jmin; j <= jmax; ++j) {
    const int b = g.AC(imin - 1, j), e = g.AC(imax, j);
    if (nobs_use) {
      for (int n = b; n < e; ++n) {
        if ((size_t)nn >= nobs_use->size()) nobs_use->resize(nn + 1024);
        (*nobs_use)[nn++] = n;
      }
    } else {
      nn += e - b;
    }
  }
}
// obs_local_range, letkf_tools.f90:1765-1788
void obs_local_range(const oracle_state &s, int ic, double ri, double rj, int &imin, int &imax,
                     int &jmin, int &jmax) {
  const double dist_zero_i = s.hori_loc_ctype[ic] * s.cfg.dist_zero_fac / s.cfg.DX;
  const double dist_zero_j = s.hori_loc_ctype[ic] * s.cfg.dist_zero_fac / s.cfg.DY;
  ij_obsgrd_ext(s, ic, ri - dist_zero_i, rj - dist_zero_j, imin, jmin);
  ij_obsgrd_ext(s, ic, ri + dist_zero_i, rj + dist_zero_j, imax, jmax);
  // the reference only asserts the range under -DDEBUG (:1779-1785); clamp so that a point
  // outside the analysed plane cannot index out of the tables
  const Mesh &g = s.obsgrd[ic];
  imin = std::max(imin, 1);
  jmin = std::max(jmin, 1);
  imax = std::min(imax, g.ngrdext_i);
  jmax = std::min(jmax, g.ngrdext_j);
}
// obs_local_cal, letkf_tools.f90:1793-1906.  iob is a 0-based sorted index.
void obs_local_cal(const oracle_state &s, double ri, double rj, double rlev, double rz, int nvar,
                   int iob, int ic, double &ndist, double &nrloc, double &nrdiag) {
  nrloc = 0.0;
  nrdiag = -1.0;
  ndist = -1.0;
  const int obelm = s.elm_ctype[ic];
  const int obtyp = s.typ_ctype[ic];
  if (nvar > 0) {   // variable localisation (:1833-1848)
    nrloc = s.var_local[nvar - 1][uid_obs_varlocal(obelm) - 1];
    if (nrloc < std::numeric_limits<double>::min()) {
      nrloc = 0.0;
      return;
    }
  }
  double nd_v;   // normalised vertical distance (:1852-1866)
  const double vl = s.vert_loc_ctype[ic];
  if (vl == 0.0) {
    nd_v = 0.0;
  } else if (obelm == ID_PS) {
    nd_v = std::fabs(std::log(s.s_dat[iob]) - std::log(rlev)) / vl;
  } else if (obelm == ID_RAIN) {
    nd_v = std::fabs(std::log(s.cfg.VERT_LOCAL_RAIN_BASE) - std::log(rlev)) / vl;
  } else if (obtyp == 22) {
    nd_v = std::fabs(s.s_lev[iob] - rz) / vl;
  } else {
    nd_v = std::fabs(std::log(s.s_lev[iob]) - std::log(rlev)) / vl;
  }
  if (nd_v > s.cfg.dist_zero_fac) {   // (:1869)
    nrloc = 0.0;
    return;
  }
  const double rdx = (ri - s.s_ri[iob]) * s.cfg.DX;   // (:1876-1878)
  const double rdy = (rj - s.s_rj[iob]) * s.cfg.DY;
  const double nd_h = std::sqrt(rdx * rdx + rdy * rdy) / s.hori_loc_ctype[ic];
  if (nd_h > s.cfg.dist_zero_fac) {   // (:1881)
    nrloc = 0.0;
    return;
  }
  ndist = nd_h * nd_h + nd_v * nd_v;   // (:1888)
  if (ndist > s.cfg.dist_zero_fac_square) {   // (:1891)
    nrloc = 0.0;
    ndist = -1.0;
    return;
  }
  nrloc = nrloc * std::exp(-0.5 * ndist);   // (:1899); nvar == 0 keeps nrloc = 0 as written
  nrdiag = s.s_err[iob] * s.s_err[iob] / nrloc;   // (:1903)
}

struct LocalOut {   // what obs_local hands to letkf_core (rows of hdxf are gathered later)
  std::vector<int> iob;
  std::vector<double> rdiag, rloc;
  void push(int i, double rd, double rl) {
    iob.push_back(i);
    rdiag.push_back(rd);
    rloc.push_back(rl);
  }
  void clear() {
    iob.clear();
    rdiag.clear();
    rloc.clear();
  }
};

struct Scratch {   // per-thread work arrays of obs_local (:1347-1351, 1400-1417)
  std::vector<int> nobs_use, nobs_use2;
  std::vector<double> dist_tmp, rloc_tmp, rdiag_tmp;
  std::vector<int> touched;   // entries of rloc_tmp written in this call (reset lazily)
};

// optional outputs nobsl_t / cutd_t of obs_local (letkf_tools.f90:1342-1343): [elm_u - 1][typ - 1]
struct LocalDiag {
  int nobsl_t[LETKF_B200_NID_OBS][LETKF_B200_NOBTYPE];
  double cutd_t[LETKF_B200_NID_OBS][LETKF_B200_NOBTYPE];
  // a limited group whose last search pass found EXACTLY the limit: no quickselect ran, so the reference's cutd_t is the
  // value of the LAST SCANNED observation (which depends on the srch_q0 history), not that of the worst selected one
  int exact_hits;
};

// obs_local, letkf_tools.f90:1325-1759.  srch_q0 may be NULL (then q starts at 1).
// brute: candidates are ALL observations of the ctype (semantic definition).
void obs_local(const oracle_state &s, double ri, double rj, double rlev, double rz, int nvar,
               LocalOut &out, Scratch &w, int *srch_q0, bool brute, LocalDiag *dg = nullptr) {
  out.clear();
  if (dg) {   // (:1380-1390)
    std::memset(dg->nobsl_t, 0, sizeof(dg->nobsl_t));
    for (auto &r : dg->cutd_t)
      for (double &v : r) v = 0.0;
    dg->exact_hits = 0;
    if (s.cfg.MAX_NOBS_PER_GRID_CRITERION == 1)
      for (int ic = 0; ic < s.nctype; ++ic)
        dg->cutd_t[s.elm_u_ctype[ic] - 1][s.typ_ctype[ic] - 1] = s.hori_loc_ctype[ic] * s.cfg.dist_zero_fac;
  }
  if (s.nobstotal == 0) return;
  int maxlimit = 0;
  for (int t = 0; t < NOBTYPE; ++t) maxlimit = std::max(maxlimit, s.cfg.MAX_NOBS_PER_GRID[t]);
  const bool limited = maxlimit > 0;
  if (limited) {
    // The reference re-initialises rloc_tmp(:) = -1.0d6 over all nobstotal entries per call
    // (:1414); only touched entries are reset here -- same values, O(candidates) cost.
    if ((int)w.rloc_tmp.size() != s.nobstotal) {
      w.rloc_tmp.assign(s.nobstotal, -1.0e6);
      w.rdiag_tmp.assign(s.nobstotal, 0.0);
      w.dist_tmp.assign(s.nobstotal, 0.0);
    }
    for (int i : w.touched) w.rloc_tmp[i] = -1.0e6;
    w.touched.clear();
  }
  auto choose = [&](int ic2, int imin, int imax, int jmin, int jmax, int &nn) {
    if (brute) {
      const Mesh &g = s.obsgrd[ic2];
      const int b = g.AC(0, 1), e = b + g.tot_ext;
      for (int n = b; n < e; ++n) {
        if ((size_t)nn >= w.nobs_use.size()) w.nobs_use.resize(nn + 1024);
        w.nobs_use[nn++] = n;
      }
    } else {
      obs_choose_ext(s, ic2, imin, imax, jmin, jmax, nn, &w.nobs_use);
    }
  };

  for (int ic = 0; ic < s.nctype; ++ic) {
    if (s.n_merge[ic] == 0) {
      if (dg) dg->cutd_t[s.elm_u_ctype[ic] - 1][s.typ_ctype[ic] - 1] = 0.0;   // (:1427-1431)
      continue;
    }
    const int nobsl_max_master = s.cfg.MAX_NOBS_PER_GRID[s.typ_ctype[ic] - 1];
    const int nm = s.n_merge[ic];
    const int eu_master = s.elm_u_ctype[ic] - 1, ty_master = s.typ_ctype[ic] - 1;

    if (nobsl_max_master <= 0) {
      // no obs-number limit (:1438-1476)
      const size_t nobsl_prev = out.iob.size();   // (:1444: set once, before the loop over the merged types)
      for (int icm = 0; icm < nm; ++icm) {
        const int ic2 = s.ic_merge[ic][icm];
        if (s.obsgrd[ic2].tot_ext > 0) {
          int nn = 0, imin, imax, jmin, jmax;
          obs_local_range(s, ic2, ri, rj, imin, imax, jmin, jmax);
          choose(ic2, imin, imax, jmin, jmax, nn);
          for (int n = 0; n < nn; ++n) {
            const int iob = w.nobs_use[n];
            double nd, nrloc, nrdiag;
            obs_local_cal(s, ri, rj, rlev, rz, nvar, iob, ic2, nd, nrloc, nrdiag);
            if (nrloc == 0.0) continue;
            out.push(iob, nrdiag, nrloc);
          }
        }
        // (:1473-1475) nobsl - nobsl_prev with nobsl_prev from BEFORE the merged loop: cumulative over the merged types
        if (dg) dg->nobsl_t[s.elm_u_ctype[ic2] - 1][s.typ_ctype[ic2] - 1] = (int)(out.iob.size() - nobsl_prev);
      }
    } else if (s.cfg.MAX_NOBS_PER_GRID_CRITERION == 1) {
      // incremental search + N nearest (:1479-1660)
      int nn = 0;
      for (int icm = 0; icm < nm; ++icm) {
        const int ic2 = s.ic_merge[ic][icm];
        if (s.obsgrd[ic2].tot_ext > 0) {
          int imin, imax, jmin, jmax;
          obs_local_range(s, ic2, ri, rj, imin, imax, jmin, jmax);
          choose(ic2, imin, imax, jmin, jmax, nn);
        }
      }
      if (nn == 0) continue;
      std::vector<double> search_incr(nm), search_incr_i(nm), search_incr_j(nm);
      std::vector<int> imin_c(nm), imax_c(nm), jmin_c(nm), jmax_c(nm), nn_steps(nm + 1);
      search_incr[0] = s.hori_loc_ctype[ic] * s.cfg.dist_zero_fac / (double)N_SEARCH_INCR;
      search_incr[0] = std::max(search_incr[0], std::max(s.obsgrd[ic].grdspc_i, s.obsgrd[ic].grdspc_j));
      for (int icm = 0; icm < nm; ++icm) {
        const int ic2 = s.ic_merge[ic][icm];
        if (icm > 0) search_incr[icm] = search_incr[0] / s.hori_loc_ctype[ic] * s.hori_loc_ctype[ic2];
        search_incr_i[icm] = search_incr[icm] / s.cfg.DX;
        search_incr_j[icm] = search_incr[icm] / s.cfg.DY;
        obs_local_range(s, ic2, ri, rj, imin_c[icm], imax_c[icm], jmin_c[icm], jmax_c[icm]);
      }
      int nobsl_incr = 0;
      int q = srch_q0 ? srch_q0[ic] - 1 : 0;
      bool loop = true;
      while (loop) {
        ++q;
        nn = 0;
        bool reach_cutoff = true;
        for (int icm = 0; icm < nm; ++icm) {
          const int ic2 = s.ic_merge[ic][icm];
          nn_steps[icm] = nn;
          if (s.obsgrd[ic2].tot_ext > 0) {
            int imin, imax, jmin, jmax;
            ij_obsgrd_ext(s, ic2, ri - search_incr_i[icm] * q, rj - search_incr_j[icm] * q, imin, jmin);
            ij_obsgrd_ext(s, ic2, ri + search_incr_i[icm] * q, rj + search_incr_j[icm] * q, imax, jmax);
            if (brute || (imin <= imin_c[icm] && imax >= imax_c[icm] && jmin <= jmin_c[icm] &&
                          jmax >= jmax_c[icm])) {
              imin = imin_c[icm];
              imax = imax_c[icm];
              jmin = jmin_c[icm];
              jmax = jmax_c[icm];
            } else {
              reach_cutoff = false;
            }
            choose(ic2, imin, imax, jmin, jmax, nn);
          }
        }
        nn_steps[nm] = nn;
        if (!reach_cutoff && nn < nobsl_max_master) continue;
        if (reach_cutoff) {
          loop = false;
          if (nn == 0) break;
        }
        nobsl_incr = 0;
        const double dist_cutoff_fac = search_incr[0] * q / s.hori_loc_ctype[ic];
        const double dist_cutoff_fac_square = dist_cutoff_fac * dist_cutoff_fac;
        for (int icm = 0; icm < nm; ++icm) {
          const int ic2 = s.ic_merge[ic][icm];
          for (int n = nn_steps[icm]; n < nn_steps[icm + 1]; ++n) {
            const int iob = w.nobs_use[n];
            if (w.rloc_tmp[iob] == 0.0) continue;
            if (w.rloc_tmp[iob] < 0.0) {
              obs_local_cal(s, ri, rj, rlev, rz, nvar, iob, ic2, w.dist_tmp[iob], w.rloc_tmp[iob],
                            w.rdiag_tmp[iob]);
              w.touched.push_back(iob);
              if (w.rloc_tmp[iob] == 0.0) continue;
            }
            if (!reach_cutoff) {
              if (w.dist_tmp[iob] > dist_cutoff_fac_square) continue;
            }
            if ((size_t)nobsl_incr >= w.nobs_use2.size()) w.nobs_use2.resize(nobsl_incr + 1024);
            w.nobs_use2[nobsl_incr++] = iob + 1;   // 1-based for quickselect
          }
        }
        if (nobsl_incr >= nobsl_max_master) loop = false;
      }
      if (srch_q0) {   // (:1604-1610)
        if (q == srch_q0[ic] && nobsl_incr > nobsl_max_master * 3) {
          srch_q0[ic] = q - 1;
        } else if (q > srch_q0[ic]) {
          srch_q0[ic] = q;
        }
      }
      if (nobsl_incr == 0) continue;
      if (dg && nobsl_incr == nobsl_max_master) dg->exact_hits++;
      if (nobsl_incr > nobsl_max_master) {
        oracle_quickselect_arg(w.dist_tmp.data(), w.nobs_use2.data(), 1, nobsl_incr, nobsl_max_master);
        nobsl_incr = nobsl_max_master;
      }
      for (int n = 0; n < nobsl_incr; ++n) {
        const int iob = w.nobs_use2[n] - 1;
        out.push(iob, w.rdiag_tmp[iob], w.rloc_tmp[iob]);
      }
      if (dg) {   // (:1653-1660)
        dg->nobsl_t[eu_master][ty_master] = nobsl_incr;
        if (nobsl_incr == nobsl_max_master)
          dg->cutd_t[eu_master][ty_master] = s.hori_loc_ctype[ic] * std::sqrt(w.dist_tmp[w.nobs_use2[nobsl_incr - 1] - 1]);
      }
    } else {
      // criterion 2 / 3: select over everything inside the cut-off (:1663-1729)
      int nn = 0, nobsl_incr = 0;
      for (int icm = 0; icm < nm; ++icm) {
        const int ic2 = s.ic_merge[ic][icm];
        if (s.obsgrd[ic2].tot_ext > 0) {
          const int nn_prev = nn;
          int imin, imax, jmin, jmax;
          obs_local_range(s, ic2, ri, rj, imin, imax, jmin, jmax);
          choose(ic2, imin, imax, jmin, jmax, nn);
          for (int n = nn_prev; n < nn; ++n) {
            const int iob = w.nobs_use[n];
            double nd;
            obs_local_cal(s, ri, rj, rlev, rz, nvar, iob, ic2, nd, w.rloc_tmp[iob], w.rdiag_tmp[iob]);
            w.touched.push_back(iob);
            if (w.rloc_tmp[iob] == 0.0) continue;
            if ((size_t)nobsl_incr >= w.nobs_use2.size()) w.nobs_use2.resize(nobsl_incr + 1024);
            w.nobs_use2[nobsl_incr++] = iob + 1;
          }
        }
      }
      if (nobsl_incr == 0) continue;
      if (dg && nobsl_incr == nobsl_max_master) dg->exact_hits++;
      if (nobsl_incr > nobsl_max_master) {
        if (s.cfg.MAX_NOBS_PER_GRID_CRITERION == 2) {
          oracle_quickselect_desc_arg(w.rloc_tmp.data(), w.nobs_use2.data(), 1, nobsl_incr, nobsl_max_master);
        } else {
          oracle_quickselect_arg(w.rdiag_tmp.data(), w.nobs_use2.data(), 1, nobsl_incr, nobsl_max_master);
        }
        nobsl_incr = nobsl_max_master;
      }
      for (int n = 0; n < nobsl_incr; ++n) {
        const int iob = w.nobs_use2[n] - 1;
        out.push(iob, w.rdiag_tmp[iob], w.rloc_tmp[iob]);
      }
      if (dg) {   // (:1718-1729)
        dg->nobsl_t[eu_master][ty_master] = nobsl_incr;
        if (nobsl_incr == nobsl_max_master) {
          const int last = w.nobs_use2[nobsl_incr - 1] - 1;
          if (s.cfg.MAX_NOBS_PER_GRID_CRITERION == 2) dg->cutd_t[eu_master][ty_master] = w.rloc_tmp[last];
          else if (s.cfg.MAX_NOBS_PER_GRID_CRITERION == 3) dg->cutd_t[eu_master][ty_master] = w.rdiag_tmp[last];
        }
      }
    }
  }
}

// relax_beta, letkf_tools.f90:1911-1948 (nlong = nlon, nlatg = nlat in the single-mesh view)
double relax_beta(const oracle_state &s, double ri, double rj, double rz) {
  const letkf_b200_config &c = s.cfg;
  double beta = 1.0;
  if (s.radar_only &&
      rz > c.RADAR_ZMAX + std::max(c.VERT_LOCAL[21], c.VERT_LOCAL_RADAR_VR) * c.dist_zero_fac) {
    return 0.0;
  }
  if (c.BOUNDARY_BUFFER_WIDTH > 0.0) {
    const double dist_bdy =
        std::min(std::min(ri - c.IHALO, c.nlon + c.IHALO + 1 - ri) * c.DX,
                 std::min(rj - c.JHALO, c.nlat + c.JHALO + 1 - rj) * c.DY) /
        c.BOUNDARY_BUFFER_WIDTH;
    if (dist_bdy < 1.0) beta = std::max(dist_bdy, 0.0);
  }
  return beta;
}

void setup_var_local(oracle_state &s) {   // letkf_tools.f90:130-163
  const int nv = s.cfg.nv3d + s.cfg.nv2d;
  for (int n = 0; n < nv; ++n)
    for (int iv = 0; iv < LETKF_B200_NID_VARLOCAL; ++iv) s.var_local[n][iv] = s.cfg.VAR_LOCAL[iv][n];
  s.n2nc_max = 1;
  s.var_local_n2nc[0] = 1;
  s.var_local_n2n[0] = 1;
  for (int n = 2; n <= nv; ++n) {
    bool found = false;
    for (int i = 1; i <= s.n2nc_max; ++i) {
      // NB: the reference indexes var_local with the GROUP number var_local_n2nc(i) as if it
      // were a variable index (:146); restated as written.
      const int ref = s.var_local_n2nc[i - 1];
      double md = 0.0;
      for (int iv = 0; iv < LETKF_B200_NID_VARLOCAL; ++iv)
        md = std::max(md, std::fabs(s.var_local[ref - 1][iv] - s.var_local[n - 1][iv]));
      if (md < std::numeric_limits<double>::min()) {
        s.var_local_n2nc[n - 1] = s.var_local_n2nc[i - 1];
        s.var_local_n2n[n - 1] = s.var_local_n2n[s.var_local_n2nc[n - 1] - 1];
        found = true;
        break;
      }
    }
    if (!found) {
      ++s.n2nc_max;
      s.var_local_n2nc[n - 1] = s.n2nc_max;
      s.var_local_n2n[n - 1] = n;
    }
  }
}

}  // namespace

extern "C" {

oracle_state *oracle_create(const letkf_b200_config *cfg) {
  oracle_state *s = new oracle_state();
  s->cfg = *cfg;
  std::memset(s->ctype_elmtyp, 0, sizeof(s->ctype_elmtyp));
  setup_var_local(*s);
  return s;
}
void oracle_destroy(oracle_state *s) { delete s; }
void oracle_set_quirks(oracle_state *s, int q) { s->quirk_ij = q; }

int oracle_set_grid(oracle_state *s, int nij1, const double *rig1, const double *rjg1,
                    const double *hgt1) {
  s->nij1 = nij1;
  s->rig1.assign(rig1, rig1 + nij1);
  s->rjg1.assign(rjg1, rjg1 + nij1);
  s->hgt1.assign(hgt1, hgt1 + (size_t)nij1 * s->cfg.nlev);
  return 0;
}

int oracle_set_obs(oracle_state *s, const letkf_b200_obs *obs) {
  const letkf_b200_config &c = s->cfg;
  const int nobs = obs->nobs;
  s->nensobs = obs->nensobs;
  // ctype table (letkf_obs.f90:300-342): ascending in (typ, elm_u)
  bool ctype_use[NID_OBS][NOBTYPE];
  std::memset(ctype_use, 0, sizeof(ctype_use));
  for (int n = 0; n < nobs; ++n) {
    const int u = uid_obs(obs->elm[n]);
    if (u < 1 || obs->typ[n] < 1 || obs->typ[n] > NOBTYPE) return -1;
    ctype_use[u - 1][obs->typ[n] - 1] = true;
  }
  s->elm_ctype.clear(); s->elm_u_ctype.clear(); s->typ_ctype.clear();
  s->hori_loc_ctype.clear(); s->vert_loc_ctype.clear();
  std::memset(s->ctype_elmtyp, 0, sizeof(s->ctype_elmtyp));
  int ictype = 0;
  for (int ityp = 1; ityp <= NOBTYPE; ++ityp)
    for (int ielm_u = 1; ielm_u <= NID_OBS; ++ielm_u)
      if (ctype_use[ielm_u - 1][ityp - 1]) {
        ++ictype;
        s->ctype_elmtyp[ielm_u - 1][ityp - 1] = ictype;
        const int elm = ELEM_UID[ielm_u - 1];
        s->elm_ctype.push_back(elm);
        s->elm_u_ctype.push_back(ielm_u);
        s->typ_ctype.push_back(ityp);
        if (elm == ID_RE0) s->hori_loc_ctype.push_back(c.HORI_LOCAL_RADAR_OBSNOREF);
        else if (elm == ID_VR) s->hori_loc_ctype.push_back(c.HORI_LOCAL_RADAR_VR);
        else s->hori_loc_ctype.push_back(c.HORI_LOCAL[ityp - 1]);
        if (elm == ID_VR) s->vert_loc_ctype.push_back(c.VERT_LOCAL_RADAR_VR);
        else s->vert_loc_ctype.push_back(c.VERT_LOCAL[ityp - 1]);
      }
  s->nctype = ictype;
  // sorting meshes (letkf_obs.f90:660-695)
  s->obsgrd.assign(s->nctype, Mesh());
  for (int ic = 0; ic < s->nctype; ++ic) {
    Mesh &g = s->obsgrd[ic];
    const int ityp = s->typ_ctype[ic];
    double target_grdspc;
    if (c.OBS_SORT_GRID_SPACING[ityp - 1] > 0) {
      target_grdspc = c.OBS_SORT_GRID_SPACING[ityp - 1];
    } else if (c.MAX_NOBS_PER_GRID[ityp - 1] > 0) {
      target_grdspc = 0.1 * std::sqrt((double)c.MAX_NOBS_PER_GRID[ityp - 1]) * c.OBS_MIN_SPACING[ityp - 1];
    } else {
      target_grdspc = s->hori_loc_ctype[ic] * c.dist_zero_fac / 6.0;
    }
    g.ngrd_i = std::min((int)std::ceil(c.DX * (double)c.nlon / target_grdspc), c.nlon);
    g.ngrd_j = std::min((int)std::ceil(c.DY * (double)c.nlat / target_grdspc), c.nlat);
    g.grdspc_i = c.DX * (double)c.nlon / (double)g.ngrd_i;
    g.grdspc_j = c.DY * (double)c.nlat / (double)g.ngrd_j;
    g.ngrdsch_i = (int)std::ceil(s->hori_loc_ctype[ic] * c.dist_zero_fac / g.grdspc_i);
    g.ngrdsch_j = (int)std::ceil(s->hori_loc_ctype[ic] * c.dist_zero_fac / g.grdspc_j);
    g.ngrdext_i = g.ngrd_i + g.ngrdsch_i * 2;
    g.ngrdext_j = g.ngrd_j + g.ngrdsch_j * 2;
    g.n_ext.assign((size_t)g.ngrdext_i * g.ngrdext_j, 0);
    g.ac_ext.assign((size_t)(g.ngrdext_i + 1) * g.ngrdext_j, 0);
    g.tot_ext = 0;
  }
  // first scan: counts per bucket (letkf_obs.f90:747-765), placed in the extended mesh (:922-941)
  std::vector<int> ob_ic(nobs), ob_i(nobs), ob_j(nobs);
  for (int n = 0; n < nobs; ++n) {
    const int ic = s->ctype_elmtyp[uid_obs(obs->elm[n]) - 1][obs->typ[n] - 1] - 1;
    int i, j;
    ij_obsgrd(*s, ic, obs->ri[n], obs->rj[n], i, j);
    Mesh &g = s->obsgrd[ic];
    i = std::min(std::max(i, 1), g.ngrd_i);
    j = std::min(std::max(j, 1), g.ngrd_j);
    ob_ic[n] = ic;
    ob_i[n] = i + g.ngrdsch_i;
    ob_j[n] = j + g.ngrdsch_j;
    g.N(ob_i[n], ob_j[n]) += 1;
  }
  // prefix sums chained across ctypes (letkf_obs.f90:943-960)
  for (int ic = 0; ic < s->nctype; ++ic) {
    Mesh &g = s->obsgrd[ic];
    if (ic > 0) {
      const Mesh &p = s->obsgrd[ic - 1];
      g.ACw(0, 1) = p.AC(p.ngrdext_i, p.ngrdext_j);
    }
    for (int j = 1; j <= g.ngrdext_j; ++j) {
      if (j > 1) g.ACw(0, j) = g.AC(g.ngrdext_i, j - 1);
      for (int i = 1; i <= g.ngrdext_i; ++i) g.ACw(i, j) = g.AC(i - 1, j) + g.N(i, j);
    }
    g.tot_ext = g.AC(g.ngrdext_i, g.ngrdext_j) - g.AC(0, 1);
  }
  s->nobstotal = s->nctype > 0 ? s->obsgrd.back().AC(s->obsgrd.back().ngrdext_i, s->obsgrd.back().ngrdext_j) : 0;
  // second scan: stable placement (letkf_obs.f90:787-805) -> order of obsda_sort
  std::vector<std::vector<int>> next(s->nctype);
  for (int ic = 0; ic < s->nctype; ++ic) {
    Mesh &g = s->obsgrd[ic];
    next[ic].resize((size_t)g.ngrdext_i * g.ngrdext_j);
    for (int j = 1; j <= g.ngrdext_j; ++j)
      for (int i = 1; i <= g.ngrdext_i; ++i) next[ic][(i - 1) + (size_t)(j - 1) * g.ngrdext_i] = g.AC(i - 1, j);
  }
  s->sorted_to_orig.assign(s->nobstotal, -1);
  for (int n = 0; n < nobs; ++n) {
    Mesh &g = s->obsgrd[ob_ic[n]];
    int &nx = next[ob_ic[n]][(ob_i[n] - 1) + (size_t)(ob_j[n] - 1) * g.ngrdext_i];
    s->sorted_to_orig[nx++] = n;
  }
  const int nt = s->nobstotal, ne = s->nensobs;
  s->s_elm.resize(nt); s->s_typ.resize(nt); s->s_ri.resize(nt); s->s_rj.resize(nt);
  s->s_lev.resize(nt); s->s_dat.resize(nt); s->s_err.resize(nt); s->s_val.resize(nt);
  s->s_ensval.resize((size_t)nt * ne);
  for (int k = 0; k < nt; ++k) {
    const int n = s->sorted_to_orig[k];
    s->s_elm[k] = obs->elm[n]; s->s_typ[k] = obs->typ[n];
    s->s_ri[k] = obs->ri[n]; s->s_rj[k] = obs->rj[n]; s->s_lev[k] = obs->lev[n];
    s->s_dat[k] = obs->dat[n]; s->s_err[k] = obs->err[n]; s->s_val[k] = obs->val[n];
    std::copy(obs->ensval + (size_t)n * ne, obs->ensval + (size_t)(n + 1) * ne,
              s->s_ensval.begin() + (size_t)k * ne);
  }
  // merged obs-number budgets (letkf_tools.f90:167-192): REF (uid 9) + RE0 (uid 10) of type 22
  int ctype_merge[NID_OBS][NOBTYPE];
  std::memset(ctype_merge, 0, sizeof(ctype_merge));
  ctype_merge[uid_obs(ID_REF) - 1][21] = 1;
  ctype_merge[uid_obs(ID_RE0) - 1][21] = 1;
  s->n_merge.assign(s->nctype, 1);
  s->ic_merge.assign(s->nctype, std::vector<int>());
  for (int ic = 0; ic < s->nctype; ++ic) {
    if (s->n_merge[ic] > 0) {
      s->ic_merge[ic].push_back(ic);
      const int cm = ctype_merge[s->elm_u_ctype[ic] - 1][s->typ_ctype[ic] - 1];
      if (cm > 0) {
        for (int ic2 = ic + 1; ic2 < s->nctype; ++ic2) {
          if (ctype_merge[s->elm_u_ctype[ic2] - 1][s->typ_ctype[ic2] - 1] == cm) {
            s->n_merge[ic] += 1;
            s->ic_merge[ic].push_back(ic2);
            s->n_merge[ic2] = 0;
          }
        }
      }
    }
  }
  s->n_merge_max = 1;
  for (int v : s->n_merge) s->n_merge_max = std::max(s->n_merge_max, v);
  s->radar_only = true;   // letkf_tools.f90:197-203
  for (int ic = 0; ic < s->nctype; ++ic)
    if (s->typ_ctype[ic] != 22) {
      s->radar_only = false;
      break;
    }
  return 0;
}

int oracle_obs_info(const oracle_state *s, int32_t *nobstotal, int32_t *nctype) {
  *nobstotal = s->nobstotal;
  *nctype = s->nctype;
  return 0;
}
int oracle_get_ctype(const oracle_state *s, int ic, letkf_b200_ctype_info *o) {
  if (ic < 0 || ic >= s->nctype) return -1;
  const Mesh &g = s->obsgrd[ic];
  o->elm = s->elm_ctype[ic]; o->elm_u = s->elm_u_ctype[ic]; o->typ = s->typ_ctype[ic];
  o->ngrd_i = g.ngrd_i; o->ngrd_j = g.ngrd_j; o->ngrdsch_i = g.ngrdsch_i; o->ngrdsch_j = g.ngrdsch_j;
  o->ngrdext_i = g.ngrdext_i; o->ngrdext_j = g.ngrdext_j; o->tot_ext = g.tot_ext;
  o->ac_begin = g.AC(0, 1); o->n_merge = s->n_merge[ic];
  o->hori_loc = s->hori_loc_ctype[ic]; o->vert_loc = s->vert_loc_ctype[ic];
  o->grdspc_i = g.grdspc_i; o->grdspc_j = g.grdspc_j;
  return 0;
}
int oracle_get_ac_ext(const oracle_state *s, int ic, int32_t *ac) {
  if (ic < 0 || ic >= s->nctype) return -1;
  std::copy(s->obsgrd[ic].ac_ext.begin(), s->obsgrd[ic].ac_ext.end(), ac);
  return 0;
}
int oracle_get_sorted_index(const oracle_state *s, int32_t *out) {
  std::copy(s->sorted_to_orig.begin(), s->sorted_to_orig.end(), out);
  return 0;
}

int oracle_obs_local(oracle_state *s, int npts, const double *ri, const double *rj,
                     const double *rlev, const double *rz, int nvar, int32_t *nobsl, int32_t *idx,
                     double *rdiag, double *rloc, int max_out, int brute) {
  int status = 0;
#ifdef _OPENMP
#pragma omp parallel
#endif
  {
    LocalOut out;
    Scratch w;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
    for (int i = 0; i < npts; ++i) {
      obs_local(*s, ri[i], rj[i], rlev[i], rz[i], nvar, out, w, nullptr, brute != 0);
      const int n = (int)out.iob.size();
      nobsl[i] = n;
      if (idx) {
        if (n > max_out) {
          status = -1;
          continue;
        }
        for (int j = 0; j < n; ++j) {
          idx[(size_t)i * max_out + j] = out.iob[j];
          if (rdiag) rdiag[(size_t)i * max_out + j] = out.rdiag[j];
          if (rloc) rloc[(size_t)i * max_out + j] = out.rloc[j];
        }
      }
    }
  }
  return status;
}

// NOBS_OUT fields of das_letkf for the 3-D model variable nvar (the reference writes those of iv3d_t): work3dn
// (letkf_tools.f90:281-284 zero-initialised, :399-401 copied inside a variable-localisation group, :440-447 filled from
// nobsl_t / cutd_t of obs_local) and the eleven fields the reference writes to NOBS_OUT_BASENAME (:767-778):
//   out(:,:,1..5)  = sum over elements of nobsl_t(:, type) for report types 1, 3, 4, 8, 22
//   out(:,:,6..8)  = nobsl_t(REF | RE0 | VR, PHARAD);   out(:,:,9..11) = cutd_t(REF | RE0 | VR, PHARAD)
// Levels run sequentially per column (search_q0 carry, :194-195, :313); points with relax_beta = 0 keep the zeros.
// pmean (nij1, nlev) = gues3d(:,:,mmean,iv3d_p).  exact_hits (nij1, nlev), optional: number of obs-number-limited groups whose
// last search pass found exactly the limit (see LocalDiag).
void oracle_nobs_out(oracle_state *sp, int nvar, const double *pmean, double *out, int32_t *exact_hits, int nthreads) {
  oracle_state &s = *sp;
  const int nij1 = s.nij1, nlev = s.cfg.nlev;
  const size_t sl = (size_t)nij1 * nlev;
  std::fill(out, out + sl * 11, 0.0);
  if (exact_hits) std::fill(exact_hits, exact_hits + sl, 0);
  const int nrep = s.var_local_n2n[nvar - 1];   // the variable whose obs_local call fills the group's work3dn
  std::vector<int> search_q0((size_t)std::max(s.nctype, 1) * nij1, 1);
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
  static const int rep_types[5] = {1, 3, 4, 8, 22};
  const int eu[3] = {9, 10, 11};   // uid_obs of id_radar_ref_obs, id_radar_ref_zero_obs, id_radar_vr_obs
#pragma omp parallel num_threads(nthreads)
  {
    LocalOut lo;
    Scratch w;
    LocalDiag dg;
    for (int il = 0; il < nlev; ++il) {
#pragma omp for schedule(dynamic, 4)
      for (int ij = 0; ij < nij1; ++ij) {
        const size_t p = ij + (size_t)il * nij1;
        if (relax_beta(s, s.rig1[ij], s.rjg1[ij], s.hgt1[p]) == 0.0) continue;   // (:323-352)
        obs_local(s, s.rig1[ij], s.rjg1[ij], pmean[p], s.hgt1[p], nrep, lo, w, &search_q0[(size_t)ij * std::max(s.nctype, 1)],
                  false, &dg);
        for (int f = 0; f < 5; ++f) {
          int sum = 0;
          for (int e = 0; e < LETKF_B200_NID_OBS; ++e) sum += dg.nobsl_t[e][rep_types[f] - 1];
          out[p + sl * f] = (double)sum;
        }
        for (int f = 0; f < 3; ++f) {
          out[p + sl * (5 + f)] = (double)dg.nobsl_t[eu[f] - 1][21];
          out[p + sl * (8 + f)] = dg.cutd_t[eu[f] - 1][21];
        }
        if (exact_hits) exact_hits[p] = dg.exact_hits;
      }
    }
  }
}

// das_letkf, letkf_tools.f90:50-932 (live path).  nv2d variables follow :528-660.
int oracle_das_letkf(oracle_state *sp, double *gues3d, double *gues2d, double *anal3d,
                     double *anal2d, double *infl3d, double *rtps_infl_out, int32_t *nobsl_out,
                     const uint8_t *point_mask, int nthreads, int64_t *npoints_out,
                     int64_t *nsolved_out) {
  oracle_state &s = *sp;
  const letkf_b200_config &c = s.cfg;
  const int k = c.MEMBER, nij1 = s.nij1, nlev = c.nlev, nv3d = c.nv3d, nv2d = c.nv2d;
  const int mmean = k + 1, mmdet = k + 2, nens = c.DET_RUN ? k + 2 : k + 1;
  const int mmdetobs = k + 1;
  if (c.DET_RUN && s.nensobs < mmdetobs && s.nobstotal > 0) return -1;
  const size_t sl = (size_t)nij1 * nlev;   // member stride
#define G3(ij, il, m, n) gues3d[(ij) + (size_t)(il)*nij1 + ((size_t)((m)-1) + (size_t)(n)*nens) * sl]
#define A3(ij, il, m, n) anal3d[(ij) + (size_t)(il)*nij1 + ((size_t)((m)-1) + (size_t)(n)*nens) * sl]
#define G2(ij, m, n) gues2d[(ij) + ((size_t)((m)-1) + (size_t)(n)*nens) * nij1]
#define A2(ij, m, n) anal2d[(ij) + ((size_t)((m)-1) + (size_t)(n)*nens) * nij1]
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
  // forecast perturbations (:209-230); skipped points keep their input when a mask is given
  if (!point_mask) {
#pragma omp parallel for collapse(2) schedule(static) num_threads(nthreads)
    for (int n = 0; n < nv3d; ++n)
      for (int m = 1; m <= k; ++m)
        for (int il = 0; il < nlev; ++il)
          for (int ij = 0; ij < nij1; ++ij) G3(ij, il, m, n) -= G3(ij, il, mmean, n);
    for (int n = 0; n < nv2d; ++n)
      for (int m = 1; m <= k; ++m)
        for (int ij = 0; ij < nij1; ++ij) G2(ij, m, n) -= G2(ij, mmean, n);
  } else {
    for (int il = 0; il < nlev; ++il)
      for (int ij = 0; ij < nij1; ++ij)
        if (point_mask[ij + (size_t)il * nij1]) {
          for (int n = 0; n < nv3d; ++n)
            for (int m = 1; m <= k; ++m) G3(ij, il, m, n) -= G3(ij, il, mmean, n);
          if (il == 0)
            for (int n = 0; n < nv2d; ++n)
              for (int m = 1; m <= k; ++m) G2(ij, m, n) -= G2(ij, mmean, n);
        }
  }
  // multiplicative inflation field work3d (:237-267)
  std::vector<double> work3d_own, work2d((size_t)nij1 * std::max(nv2d, 1), c.INFL_MUL);
  double *work3d = infl3d;
  if (!work3d || c.INFL_MUL > 0.0) {
    if (!work3d) {
      work3d_own.assign(sl * nv3d, c.INFL_MUL);
      work3d = work3d_own.data();
    } else {
      std::fill(work3d, work3d + sl * nv3d, c.INFL_MUL);
    }
  }
  if (c.INFL_MUL_MIN > 0.0) {
    for (size_t i = 0; i < sl * nv3d; ++i) work3d[i] = std::max(work3d[i], c.INFL_MUL_MIN);
    for (double &v : work2d) v = std::max(v, c.INFL_MUL_MIN);
  }
  if (rtps_infl_out) std::fill(rtps_infl_out, rtps_infl_out + sl * nv3d, 1.0);
  if (nobsl_out) std::fill(nobsl_out, nobsl_out + sl, 0);

  std::vector<int> search_q0((size_t)std::max(s.nctype, 1) * (nv3d + 1) * nij1, 1);   // (:194-195)
  int status = 0;
  int64_t npoints = 0, nsolved = 0;
  const int ngroups = s.n2nc_max;

#pragma omp parallel num_threads(nthreads) reduction(+ : npoints, nsolved)
  {
    LocalOut lo;
    Scratch w;
    std::vector<double> hdxf, dep, depd, trans((size_t)k * k * ngroups), transm((size_t)k * ngroups),
        transmd((size_t)k * ngroups), pa((size_t)k * k * ngroups), transrlx((size_t)k * k), q_anal(k), xb(k);
    std::vector<char> trans_done(nv3d + nv2d + 1);
    std::vector<int> nobsl_g(ngroups);
    for (int il = 0; il < nlev; ++il) {   // levels are sequential (search_q0 carry, :313)
#pragma omp for schedule(dynamic, 4)
      for (int ij = 0; ij < nij1; ++ij) {
        if (point_mask && !point_mask[ij + (size_t)il * nij1]) continue;
        ++npoints;
        std::fill(trans_done.begin(), trans_done.end(), 0);
        const double beta = relax_beta(s, s.rig1[ij], s.rjg1[ij], s.hgt1[ij + (size_t)il * nij1]);
        if (beta == 0.0) {   // (:333-359)
          for (int n = 0; n < nv3d; ++n) {
            for (int m = 1; m <= k; ++m) A3(ij, il, m, n) = G3(ij, il, mmean, n) + G3(ij, il, m, n);
            if (c.DET_RUN) A3(ij, il, mmdet, n) = G3(ij, il, mmdet, n);
          }
          if (il == 0)
            for (int n = 0; n < nv2d; ++n) {
              for (int m = 1; m <= k; ++m) A2(ij, m, n) = G2(ij, mmean, n) + G2(ij, m, n);
              if (c.DET_RUN) A2(ij, mmdet, n) = G2(ij, mmdet, n);
            }
          continue;
        }
        const double pmean = G3(ij, il, mmean, c.iv3d_p - 1);
        const int nvtot = nv3d + (il == 0 ? nv2d : 0);
        bool solved_any = false;
        for (int nn = 1; nn <= nvtot; ++nn) {
          const bool is2d = nn > nv3d;
          const int n = is2d ? nn - nv3d : nn;   // 1-based index inside its family
          const int n2nc = s.var_local_n2nc[nn - 1];
          const int n2n = s.var_local_n2n[nn - 1];
          if (!is2d && pmean < c.Q_UPDATE_TOP && n >= c.iv3d_q && n <= c.iv3d_qg) {   // (:371-385)
            for (int m = 1; m <= k; ++m) A3(ij, il, m, n - 1) = G3(ij, il, mmean, n - 1) + G3(ij, il, m, n - 1);
            if (c.DET_RUN) A3(ij, il, mmdet, n - 1) = G3(ij, il, mmdet, n - 1);
            continue;
          }
          double *infl_slot = is2d ? &work2d[ij + (size_t)(n - 1) * nij1]
                                   : &work3d[ij + (size_t)il * nij1 + (size_t)(n - 1) * sl];
          const double parm = c.RELAX_TO_INFLATED_PRIOR ? *infl_slot : 1.0;   // (:387-391)
          double *tr = &trans[(size_t)(n2nc - 1) * k * k], *trm = &transm[(size_t)(n2nc - 1) * k];
          double *trmd = &transmd[(size_t)(n2nc - 1) * k], *pag = &pa[(size_t)(n2nc - 1) * k * k];
          if (trans_done[n2nc]) {
            if (c.INFL_MUL_ADAPTIVE) {   // (:396-398, 545-556)
              *infl_slot = (n2n <= nv3d) ? work3d[ij + (size_t)il * nij1 + (size_t)(n2n - 1) * sl]
                                         : work2d[ij + (size_t)(n2n - nv3d - 1) * nij1];
            }
          } else {
            int *sq0 = &search_q0[((size_t)ij * (nv3d + 1) + (is2d ? nv3d : n - 1)) * std::max(s.nctype, 1)];
            obs_local(s, s.rig1[ij], s.rjg1[ij], pmean, s.hgt1[ij + (size_t)il * nij1], nn, lo, w, sq0, false);
            const int nobsl = (int)lo.iob.size();
            nobsl_g[n2nc - 1] = nobsl;
            const int ld = std::max(nobsl, 1);
            hdxf.resize((size_t)ld * k);
            dep.resize(ld);
            depd.resize(ld);
            for (int i = 0; i < nobsl; ++i) {   // gather rows of ensval (:1461-1469)
              const double *ev = &s.s_ensval[(size_t)lo.iob[i] * s.nensobs];
              for (int m = 0; m < k; ++m) hdxf[i + (size_t)m * ld] = ev[m];
              dep[i] = s.s_val[lo.iob[i]];
              if (c.DET_RUN) depd[i] = ev[mmdetobs - 1];
            }
            const bool want_pa = c.RELAX_ALPHA_SPREAD != 0.0;
            int r = oracle_letkf_core(k, ld, nobsl, hdxf.data(), lo.rdiag.data(), lo.rloc.data(), dep.data(),
                                      infl_slot, tr, trm, want_pa ? pag : nullptr, 1, c.INFL_MUL_ADAPTIVE,
                                      c.DET_RUN ? depd.data() : nullptr, c.DET_RUN ? trmd : nullptr);
            if (r != 0) {
#pragma omp critical
              status = LETKF_B200_EEIGEN;
            }
            trans_done[n2nc] = 1;
            if (nobsl > 0) solved_any = true;
            if (nobsl_out && n2nc == 1 && !is2d) nobsl_out[ij + (size_t)il * nij1] = nobsl;
          }
          // relaxation (:457-469, 1953-2002)
          auto X = [&](int m) -> double { return is2d ? G2(ij, m, n - 1) : G3(ij, il, m, n - 1); };
          if (c.RELAX_ALPHA != 0.0) {
            for (size_t i = 0; i < (size_t)k * k; ++i) transrlx[i] = (1.0 - c.RELAX_ALPHA) * tr[i];
            for (int m = 0; m < k; ++m) transrlx[m + (size_t)m * k] += c.RELAX_ALPHA * std::sqrt(parm);
          } else if (c.RELAX_ALPHA_SPREAD != 0.0) {
            double var_g = 0.0, var_a = 0.0;
            for (int m = 1; m <= k; ++m) xb[m - 1] = X(m);
            for (int m = 0; m < k; ++m) {
              var_g += xb[m] * xb[m];
              for (int kk = 0; kk < k; ++kk) var_a += xb[kk] * pag[kk + (size_t)m * k] * xb[m];
            }
            double infl_out = 1.0;
            if (var_g > 0.0 && var_a > 0.0) {
              infl_out = c.RELAX_ALPHA_SPREAD * std::sqrt(var_g * parm / (var_a * (double)(k - 1))) -
                         c.RELAX_ALPHA_SPREAD + 1.0;
              for (size_t i = 0; i < (size_t)k * k; ++i) transrlx[i] = tr[i] * infl_out;
            } else {
              std::copy(tr, tr + (size_t)k * k, transrlx.begin());
            }
            if (rtps_infl_out && !is2d) rtps_infl_out[ij + (size_t)il * nij1 + (size_t)(n - 1) * sl] = infl_out;
          } else {
            std::copy(tr, tr + (size_t)k * k, transrlx.begin());
          }
          // total weight matrix (:472-477)
          for (int m = 0; m < k; ++m) {
            for (int kk = 0; kk < k; ++kk) transrlx[kk + (size_t)m * k] = (transrlx[kk + (size_t)m * k] + trm[kk]) * beta;
            transrlx[m + (size_t)m * k] += (1.0 - beta);
          }
          // analysis update (:480-497)
          for (int m = 1; m <= k; ++m) {
            double a = is2d ? G2(ij, mmean, n - 1) : G3(ij, il, mmean, n - 1);
            for (int kk = 1; kk <= k; ++kk) a = a + X(kk) * transrlx[(kk - 1) + (size_t)(m - 1) * k];
            if (is2d) A2(ij, m, n - 1) = a; else A3(ij, il, m, n - 1) = a;
          }
          if (c.DET_RUN) {
            double a = 0.0;
            for (int kk = 1; kk <= k; ++kk) a = a + X(kk) * trmd[kk - 1];
            if (is2d) A2(ij, mmdet, n - 1) = G2(ij, mmdet, n - 1) + a * beta;
            else A3(ij, il, mmdet, n - 1) = G3(ij, il, mmdet, n - 1) + a * beta;
          }
          // limit q spread (:500-513)
          if (!is2d && c.Q_SPRD_MAX > 0.0 && n == c.iv3d_q) {
            double q_mean = 0.0;
            for (int m = 1; m <= k; ++m) q_mean += A3(ij, il, m, n - 1);
            q_mean /= (double)k;
            double q_sprd = 0.0;
            for (int m = 1; m <= k; ++m) {
              q_anal[m - 1] = A3(ij, il, m, n - 1) - q_mean;
              q_sprd += q_anal[m - 1] * q_anal[m - 1];
            }
            q_sprd = std::sqrt(q_sprd / (double)(k - 1)) / q_mean;
            if (q_sprd > c.Q_SPRD_MAX)
              for (int m = 1; m <= k; ++m) A3(ij, il, m, n - 1) = q_mean + q_anal[m - 1] * c.Q_SPRD_MAX / q_sprd;
          }
        }
        if (solved_any) ++nsolved;
      }
    }
  }
  if (npoints_out) *npoints_out = npoints;
  if (nsolved_out) *nsolved_out = nsolved;
  return status;
#undef G3
#undef A3
#undef G2
#undef A2
}

// ensmean_grd, common_scale.f90:1513-1552: sequential sum m = 1..mem, then divide.
void oracle_ensmean_grd(int mem, int nens, int nij, int nlev, int nv3d, int nv2d, double *v3d,
                        double *v2d) {
  const size_t sl = (size_t)nij * nlev;
  for (int n = 0; n < nv3d; ++n)
    for (size_t p = 0; p < sl; ++p) {
      double *b = v3d + p + (size_t)n * nens * sl;
      double a = b[0];
      for (int m = 1; m < mem; ++m) a += b[(size_t)m * sl];
      b[(size_t)mem * sl] = a / (double)mem;
    }
  for (int n = 0; n < nv2d; ++n)
    for (int i = 0; i < nij; ++i) {
      double *b = v2d + i + (size_t)n * nens * nij;
      double a = b[0];
      for (int m = 1; m < mem; ++m) a += b[(size_t)m * nij];
      b[(size_t)mem * nij] = a / (double)mem;
    }
}

// set_letkf_obs, letkf_obs.f90:355-560: radar acceptance, ensemble mean / perturbation of H(x),
// departures, gross-error QC.  qc codes: common_obs_scale.f90:139-151.
void oracle_obs_departure_qc(const letkf_b200_qc_config *q, int member, int det, int nobs, int nensobs,
                             const int32_t *elm, const double *dat, const double *err, int32_t *qc,
                             double *ensval, double *val) {
  const int iqc_gross_err = 5, iqc_ref_mem = 12, iqc_obs_bad = 50, iqc_otype = 90;
  const double undef = -9.99e33;   // common/common.f90:38
  auto ge = [&](double v) { return v < 0.0 ? q->GROSS_ERROR : v; };   // common_nml.f90:619-642
  for (int n = 0; n < nobs; ++n) {
    if (qc[n] > 0) continue;   // :362
    double *ev = ensval + (size_t)n * nensobs;
    if (elm[n] == ID_REF || elm[n] == ID_RE0) {   // :370-414
      if (!q->USE_RADAR_REF) { qc[n] = iqc_otype; continue; }
      if (dat[n] == undef) { qc[n] = iqc_obs_bad; continue; }
      int mem_ref = 0;
      for (int i = 0; i < member; ++i)
        if (ev[i] > q->RADAR_REF_THRES_DBZ + 1.0e-6) ++mem_ref;
      if (dat[n] > q->RADAR_REF_THRES_DBZ + 1.0e-6) {
        if (mem_ref < q->MIN_RADAR_REF_MEMBER_OBSREF) { qc[n] = iqc_ref_mem; continue; }
      } else {
        if (mem_ref < q->MIN_RADAR_REF_MEMBER) { qc[n] = iqc_ref_mem; continue; }
      }
    }
    if (elm[n] == ID_VR && !q->USE_RADAR_VR) { qc[n] = iqc_otype; continue; }   // :416-421
    double v = ev[0];   // :474-478
    for (int i = 1; i < member; ++i) v = v + ev[i];
    v = v / (double)member;
    for (int i = 0; i < member; ++i) ev[i] = ev[i] - v;   // :486-488
    val[n] = dat[n] - v;                                  // :489
    if (det) ev[member] = dat[n] - ev[member];            // :490-492
    double g;
    switch (elm[n]) {   // :503-549
      case ID_RAIN: g = ge(q->GROSS_ERROR_RAIN); break;
      case ID_REF: case ID_RE0: g = ge(q->GROSS_ERROR_RADAR_REF); break;
      case ID_VR: g = ge(q->GROSS_ERROR_RADAR_VR); break;
      case 4003: g = ge(q->GROSS_ERROR_RADAR_PRH); break;
      case 99991: g = ge(q->GROSS_ERROR_TCX); break;
      case 99992: g = ge(q->GROSS_ERROR_TCY); break;
      case 99993: g = ge(q->GROSS_ERROR_TCP); break;
      default: g = q->GROSS_ERROR; break;
    }
    if (std::fabs(val[n]) > g * err[n]) qc[n] = iqc_gross_err;
  }
}

// enssprd_grd, common_scale.f90:1557-1611 (3-D part)
void oracle_enssprd_grd(int mem, int nens, int nij, int nlev, int nv3d, const double *v3d, double *v3ds) {
  const size_t sl = (size_t)nij * nlev;
  for (int n = 0; n < nv3d; ++n)
    for (size_t p = 0; p < sl; ++p) {
      const double *b = v3d + p + (size_t)n * nens * sl;
      const double mean = b[(size_t)mem * sl];
      double a = (b[0] - mean) * (b[0] - mean);
      for (int m = 1; m < mem; ++m) a = a + (b[(size_t)m * sl] - mean) * (b[(size_t)m * sl] - mean);
      v3ds[p + (size_t)n * sl] = std::sqrt(a / (double)(mem - 1));
    }
}

// Additive inflation block of das_letkf, letkf_tools.f90:804-929 (the ensemble arrives as an argument instead of
// read_ens_mpi_addiinfl; ishuf comes from the caller instead of Knuth_Shuffle).  addi3d/addi2d are turned into
// perturbations in place like the reference's gues3d/gues2d.  obs_ri/obs_rj: positions of the radar-reflectivity
// observations of combined type (REF, PHARAD) -- the loop :822-830 over obsgrd(ic)%ac_ext -- hloc its hori_loc_ctype.
void oracle_additive_inflation(int mem, int nens, int nij, int nlev, int nv3d, int nv2d, double infl_add, int q_ratio,
                               int ref_only, int iv3d_q, int iv3d_qg, const int32_t *ishuf, double *addi3d, double *addi2d,
                               const double *gues3d, double *anal3d, double *anal2d, const double *rig1, const double *rjg1,
                               int nref, const double *obs_ri, const double *obs_rj, double hloc, double DX, double DY,
                               double dist_zero_fac_square, double *weight) {
  const size_t sl = (size_t)nij * nlev;
  std::vector<double> w((size_t)nij, 1.0);
  if (ref_only) {   // :816-840
    for (int ij = 0; ij < nij; ++ij) {
      double ref_min_dist = 1.0e33;
      for (int o = 0; o < nref; ++o) {
        const double rdx = (rig1[ij] - obs_ri[o]) * DX, rdy = (rjg1[ij] - obs_rj[o]) * DY;
        const double rdxy = rdx * rdx + rdy * rdy;
        if (rdxy < ref_min_dist) ref_min_dist = rdxy;
      }
      ref_min_dist = ref_min_dist / (hloc * hloc);
      w[ij] = (ref_min_dist <= dist_zero_fac_square) ? std::exp(-0.5 * ref_min_dist) : 0.0;
    }
  }
  if (weight) std::copy(w.begin(), w.end(), weight);
  oracle_ensmean_grd(mem, nens, nij, nlev, nv3d, addi2d ? nv2d : 0, addi3d, addi2d);   // :849
  for (int n = 0; n < nv3d; ++n)   // :869-877
    for (int m = 0; m < mem; ++m)
      for (size_t p = 0; p < sl; ++p) addi3d[p + ((size_t)m + (size_t)n * nens) * sl] -= addi3d[p + ((size_t)mem + (size_t)n * nens) * sl];
  if (addi2d)
    for (int n = 0; n < nv2d; ++n)
      for (int m = 0; m < mem; ++m)
        for (int i = 0; i < nij; ++i) addi2d[i + ((size_t)m + (size_t)n * nens) * nij] -= addi2d[i + ((size_t)mem + (size_t)n * nens) * nij];
  for (int n = 0; n < nv3d; ++n) {   // :889-912
    const bool moist = (n + 1) >= iv3d_q && (n + 1) <= iv3d_qg;   // q, qc, qr, qi, qs, qg: contiguous variable indices
    for (int m = 0; m < mem; ++m) {
      const int ms = ishuf ? ishuf[m] - 1 : m;
      for (size_t p = 0; p < sl; ++p) {
        const double g = addi3d[p + ((size_t)ms + (size_t)n * nens) * sl];
        double &a = anal3d[p + ((size_t)m + (size_t)n * nens) * sl];
        if (moist) {
          const double work = q_ratio ? gues3d[p + ((size_t)mem + (size_t)n * nens) * sl] : 1.0;   // work3d (:807-811)
          a = a + g * infl_add * w[p % nij] * work;
        } else {
          a = a + g * infl_add * w[p % nij];
        }
      }
    }
  }
  if (addi2d && anal2d)
    for (int n = 0; n < nv2d; ++n)   // :914-925
      for (int m = 0; m < mem; ++m) {
        const int ms = ishuf ? ishuf[m] - 1 : m;
        for (int i = 0; i < nij; ++i)
          anal2d[i + ((size_t)m + (size_t)n * nens) * nij] =
              anal2d[i + ((size_t)m + (size_t)n * nens) * nij] + addi2d[i + ((size_t)ms + (size_t)n * nens) * nij] * infl_add * w[i];
      }
}

// state_trans (common_scale.f90:1181-1224) and state_trans_inv (:1229-1280), in place.
void oracle_state_trans(const letkf_b200_thermo *t, int inverse, int nlev, int nlon, int nlat, int nv3d,
                        int iv3d_q, double *v3dg) {
  const size_t st = (size_t)nlev * nlon * nlat;
  const int iq = iv3d_q - 1;
  if (inverse) {   // :1243-1250
    for (int n = 0; n < nv3d; ++n) {
      const bool clamp = (n == iq) ? t->POSITIVE_DEFINITE_Q != 0 : (t->POSITIVE_DEFINITE_QHYD != 0 && n > iq && n <= iq + 5);
      if (clamp)
        for (size_t p = 0; p < st; ++p) v3dg[n * st + p] = std::max(v3dg[n * st + p], 0.0);
    }
  }
  for (size_t p = 0; p < st; ++p) {
    double *v = v3dg + p;
    double qdry = 1.0, CVtot = 0.0;
    for (int n = iq; n < nv3d; ++n) {
      qdry = qdry - v[n * st];
      CVtot = CVtot + v[n * st] * t->TRACER_CV[n - iq];
    }
    CVtot = t->CVdry * qdry + CVtot;
    const double Rtot = t->Rdry * qdry + t->Rvap * v[iq * st];
    if (!inverse) {
      const double CPovCV = (CVtot + Rtot) / CVtot;
      const double rho = v[0];
      const double pres = t->PRE00 * std::pow(v[4 * st] * Rtot / t->PRE00, CPovCV);
      const double temp = pres / (rho * Rtot);
      v[0] = v[1 * st] / rho;
      v[1 * st] = v[2 * st] / rho;
      v[2 * st] = v[3 * st] / rho;
      v[3 * st] = temp;
      v[4 * st] = pres;
    } else {
      const double CVovCP = CVtot / (CVtot + Rtot);
      const double rho = v[4 * st] / (Rtot * v[3 * st]);
      const double rhot = t->PRE00 / Rtot * std::pow(v[4 * st] / t->PRE00, CVovCP);
      v[4 * st] = rhot;
      v[3 * st] = v[2 * st] * rho;
      v[2 * st] = v[1 * st] * rho;
      v[1 * st] = v[0] * rho;
      v[0] = rho;
    }
  }
}

// monit_dep, common_obs_scale.f90:1851-1895
void oracle_monit_dep(int nn, const int32_t *elm, const double *dep, const int32_t *qc, int32_t *nobs,
                      double *bias, double *rmse) {
  for (int i = 0; i < NID_OBS; ++i) { nobs[i] = 0; bias[i] = 0.0; rmse[i] = 0.0; }
  for (int n = 0; n < nn; ++n) {
    if (qc[n] != 0) continue;
    int ielm = elm[n];
    if (ielm == 3074) ielm = 3073;        // Tv as T
    if (ielm == ID_RE0) ielm = ID_REF;    // RE0 as REF
    const int i = uid_obs(ielm) - 1;
    if (i < 0) continue;
    nobs[i] += 1;
    bias[i] = bias[i] + dep[n];
    rmse[i] = rmse[i] + dep[n] * dep[n];
  }
  for (int i = 0; i < NID_OBS; ++i) {
    if (nobs[i] == 0) { bias[i] = -9.99e33; rmse[i] = -9.99e33; }
    else { bias[i] = bias[i] / (double)nobs[i]; rmse[i] = std::sqrt(rmse[i] / (double)nobs[i]); }
  }
}

// set_common_mpi_grid, common_mpi_scale.f90:264-283
void oracle_nij1(int nlon, int nlat, int np, int myrank_e, int32_t *nij1, int32_t *nij1max) {
  const int i = (nlon * nlat) % np;
  *nij1max = (nlon * nlat - i) / np + 1;
  *nij1 = (myrank_e < i) ? *nij1max : *nij1max - 1;
}

namespace {
const double UNDEF = -9.99e33;   // common/common.f90:38
}

// scatter_grd_mpi_alltoall pack half (:1296-1308) with grd_to_buf (:1428-1455)
void oracle_grd_to_buf(int nlon, int nlat, int nlev, int nv3d, int nv2d, int np, const double *v3dg,
                       const double *v2dg, double *bufs) {
  int32_t n1, nmax;
  oracle_nij1(nlon, nlat, np, 0, &n1, &nmax);
  const int nlevall = nlev * nv3d + nv2d;
  for (int m = 1; m <= np; ++m) {
    int32_t nij_m, dummy;
    oracle_nij1(nlon, nlat, np, m - 1, &nij_m, &dummy);
    for (int jl = 0; jl < nlevall; ++jl) {
      double *buf = bufs + (size_t)jl * nmax + (size_t)(m - 1) * nmax * nlevall;
      for (int i = 1; i <= nij_m; ++i) {
        const int j = m - 1 + np * (i - 1);
        const int ilon = j % nlon + 1;
        const int ilat = (j - ilon + 1) / nlon + 1;
        double v;
        if (jl < nlev * nv3d) {
          const int n = jl / nlev, k = jl % nlev;
          v = v3dg[k + (size_t)nlev * ((ilon - 1) + (size_t)nlon * ((ilat - 1) + (size_t)nlat * n))];
        } else {
          const int n = jl - nlev * nv3d;
          v = v2dg[(ilon - 1) + (size_t)nlon * ((ilat - 1) + (size_t)nlat * n)];
        }
        buf[i - 1] = v;
      }
      if (nij_m < nmax) buf[nmax - 1] = UNDEF;
    }
  }
}
// scatter_grd_mpi_alltoall unpack half (:1319-1332)
void oracle_buf_to_ens(int nlon, int nlat, int nlev, int nv3d, int nv2d, int np, int myrank_e,
                       int nens, int mstart, int mend, const double *bufr, double *v3d, double *v2d) {
  int32_t nij1, nmax;
  oracle_nij1(nlon, nlat, np, myrank_e, &nij1, &nmax);
  const int nlevall = nlev * nv3d + nv2d;
  for (int m = mstart; m <= mend; ++m) {
    int j = 0;
    for (int n = 0; n < nv3d; ++n)
      for (int k = 0; k < nlev; ++k, ++j)
        for (int i = 0; i < nij1; ++i)
          v3d[i + (size_t)nij1 * (k + (size_t)nlev * ((m - 1) + (size_t)nens * n))] =
              bufr[i + (size_t)nmax * (j + (size_t)nlevall * (m - mstart))];
    for (int n = 0; n < nv2d; ++n, ++j)
      for (int i = 0; i < nij1; ++i)
        v2d[i + (size_t)nij1 * ((m - 1) + (size_t)nens * n)] =
            bufr[i + (size_t)nmax * (j + (size_t)nlevall * (m - mstart))];
  }
}
// gather_grd_mpi_alltoall pack half (:1358-1369)
void oracle_ens_to_buf(int nlon, int nlat, int nlev, int nv3d, int nv2d, int np, int myrank_e,
                       int nens, int mstart, int mend, const double *v3d, const double *v2d, double *bufs) {
  int32_t nij1, nmax;
  oracle_nij1(nlon, nlat, np, myrank_e, &nij1, &nmax);
  const int nlevall = nlev * nv3d + nv2d;
  for (int m = mstart; m <= mend; ++m) {
    int j = 0;
    for (int n = 0; n < nv3d; ++n)
      for (int k = 0; k < nlev; ++k, ++j)
        for (int i = 0; i < nij1; ++i)
          bufs[i + (size_t)nmax * (j + (size_t)nlevall * (m - mstart))] =
              v3d[i + (size_t)nij1 * (k + (size_t)nlev * ((m - 1) + (size_t)nens * n))];
    for (int n = 0; n < nv2d; ++n, ++j)
      for (int i = 0; i < nij1; ++i)
        bufs[i + (size_t)nmax * (j + (size_t)nlevall * (m - mstart))] =
            v2d[i + (size_t)nij1 * ((m - 1) + (size_t)nens * n)];
  }
}
// gather_grd_mpi_alltoall unpack half (:1380-1392) with buf_to_grd (:1460-1476)
void oracle_buf_to_grd(int nlon, int nlat, int nlev, int nv3d, int nv2d, int np, const double *bufr,
                       double *v3dg, double *v2dg) {
  int32_t n1, nmax;
  oracle_nij1(nlon, nlat, np, 0, &n1, &nmax);
  const int nlevall = nlev * nv3d + nv2d;
  for (int m = 1; m <= np; ++m) {
    int32_t nij_m, dummy;
    oracle_nij1(nlon, nlat, np, m - 1, &nij_m, &dummy);
    for (int jl = 0; jl < nlevall; ++jl) {
      const double *buf = bufr + (size_t)jl * nmax + (size_t)(m - 1) * nmax * nlevall;
      for (int i = 1; i <= nij_m; ++i) {
        const int j = m - 1 + np * (i - 1);
        const int ilon = j % nlon + 1;
        const int ilat = (j - ilon + 1) / nlon + 1;
        if (jl < nlev * nv3d) {
          const int n = jl / nlev, k = jl % nlev;
          v3dg[k + (size_t)nlev * ((ilon - 1) + (size_t)nlon * ((ilat - 1) + (size_t)nlat * n))] = buf[i - 1];
        } else {
          const int n = jl - nlev * nv3d;
          v2dg[(ilon - 1) + (size_t)nlon * ((ilat - 1) + (size_t)nlat * n)] = buf[i - 1];
        }
      }
    }
  }
}

}  // extern "C"
