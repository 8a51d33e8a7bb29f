/* das_from_c.c -- calling the LETKF analysis path through the C ABI from plain C (what the ISO_C_BINDING
 * module scale_letkf_b200/fortran/letkf_b200_iface.f90 does from Fortran).  Column-major arrays, host memory.
 *
 *   gcc -std=c99 -I include examples/das_from_c.c -L scale_letkf_b200 -lletkf_b200 -o das_from_c
 *
 * The caller provides: the namelist scalars (letkf_b200_config), rig1/rjg1/hgt1 of its columns, the QC-passed
 * observations with their H(x) perturbations, and gues3d(nij1,nlev,nens,nv3d); it receives anal3d.  Replaces
 * set_letkf_obs (bucket sort) + das_letkf of scale/letkf/letkf.f90:150,196. */
#include <stdio.h>
#include <stdlib.h>

#include "letkf_b200.h"

int run_analysis(int member, int nlon, int nlat, int nlev, int nij1, const double *rig1, const double *rjg1,
                 const double *hgt1, const letkf_b200_obs *obs, double *gues3d, double *anal3d) {
  letkf_b200_config cfg;
  letkf_b200_handle *h = NULL;
  letkf_b200_das_args a;
  int64_t npoints = 0, nsolved = 0, nfail = 0, nobsl_sum = 0;
  int rc;

  letkf_b200_config_defaults(&cfg);   /* reference defaults of common_nml.f90 */
  cfg.MEMBER = member;
  cfg.nlon = nlon;
  cfg.nlat = nlat;
  cfg.nlev = nlev;
  cfg.RELAX_ALPHA_SPREAD = 0.95;      /* RTPS */
  letkf_b200_config_resolve(&cfg);

  rc = letkf_b200_create(&cfg, 0, &h);
  if (rc != LETKF_B200_OK) {
    fprintf(stderr, "letkf_b200_create: %d (no usable CUDA device? there is no CPU fallback)\n", rc);
    return rc;
  }
  rc = letkf_b200_set_grid(h, nij1, rig1, rjg1, hgt1, LETKF_B200_MEM_HOST);
  if (rc == LETKF_B200_OK) rc = letkf_b200_set_obs(h, obs);
  if (rc == LETKF_B200_OK) {
    a.gues3d = gues3d;   /* INOUT: destroyed -> perturbations, slot MEMBER+1 = mean */
    a.gues2d = NULL;
    a.anal3d = anal3d;
    a.anal2d = NULL;
    a.infl3d = NULL;
    a.rtps_infl_out = NULL;
    a.nobsl_out = NULL;
    a.logp = NULL;
    a.mem_space = LETKF_B200_MEM_HOST;
    a.reserved = 0;
    rc = letkf_b200_das_letkf(h, &a);
  }
  if (rc != LETKF_B200_OK) fprintf(stderr, "letkf_b200: %d: %s\n", rc, letkf_b200_last_error(h));
  else {
    letkf_b200_das_stats(h, &npoints, &nsolved, &nfail, &nobsl_sum);
    printf("analysed %lld points, %lld with local observations\n", (long long)npoints, (long long)nsolved);
  }
  letkf_b200_destroy(h);
  return rc;
}
