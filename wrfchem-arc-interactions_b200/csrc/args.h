// Kernel argument blocks (device pointers in WRF memory order) shared by api.cu and the kernels.
#pragma once
#include <stdint.h>

#include "common.cuh"

namespace arc {

struct Geo {
  int ims, ime, kms, kme, jms, jme, its, ite, jts, jte, kts, kte;
  int ni, nk;      // memory extents in i and k
  int nci;         // tile columns per row (ite-its+1)
  int ncol_tile;   // nci * (jte-jts+1)
  __host__ __device__ size_t at3(int i, int k, int j) const { return (size_t)(i - ims) + (size_t)ni * ((size_t)(k - kms) + (size_t)nk * (size_t)(j - jms)); }
  __host__ __device__ size_t at2(int i, int j) const { return (size_t)(i - ims) + (size_t)ni * (size_t)(j - jms); }
  __host__ __device__ size_t atp(int i, int k, int j) const { return (size_t)(i - ims) + (size_t)ni * ((size_t)(k - kms) + (size_t)(nk + 2) * (size_t)(j - jms)); }
  __host__ __device__ size_t n3() const { return (size_t)ni * nk * (size_t)(jme - jms + 1); }
};

struct DebugTaps {     // device pointers (nullable)
  int *laytrop, *jp, *jt, *jt1, *indfor, *indself, *indminor;
  float *fac00, *fac01, *fac10, *fac11;
  unsigned char *cldmask;
  float *taug, *taur, *sfluxzen, *taucmc, *hr;
};

// inputs common to the SW and LW adapters
struct CloudFields {
  int icloud, warm_rain, is_cammgmp_used, has_reqc, has_reqi, has_reqs, progn;
  int f_qv, f_qc, f_qr, f_qi, f_qs, f_qg, f_qndrop;
  float g;
  const float *t3d, *cldfra3d, *lradius, *iradius, *qv3d, *qc3d, *qr3d, *qi3d, *qs3d, *qg3d, *qndrop3d;
  const float *re_cloud, *re_ice, *re_snow, *f_ice_phy, *xland, *xice, *snow;
};

struct SwArgs {
  Geo geo;
  DevTables tb;
  CloudFields cf;
  int nlay;                 // kte-kts+2
  int ncols;                // sunlit columns to process
  const int *cols;          // tile column ids (c = (j-jts)*nci + (i-its)) of the sunlit columns
  const uint32_t *mask;     // McICA bits [(g*W + w)*ncols + ci]
  int W;
  int variants;             // ARC_VAR_* mask
  int o3input, aer_ra_feedback, sf_surface_physics;
  float solcon;
  const float *t8w, *p3d, *p8w, *pi3d, *o33d, *tsk;
  const float *tauaer300, *tauaer400, *tauaer600, *tauaer999, *gaer400, *gaer600, *waer400, *waer600;
  const float *tauaer3d_sw, *ssaaer3d_sw, *asyaer3d_sw;
  const float *xcoszen, *albedo, *alswvisdir, *alswvisdif, *alswnirdir, *alswnirdif;
  // outputs
  float *rthratensw, *gsw, *swcf;
  float *swupt, *swuptc, *swuptcln, *swdnt, *swdntc, *swdntcln, *swupb, *swupbc, *swupbcln, *swdnb, *swdnbc, *swdnbcln;
  float *swvisdir, *swvisdif, *swnirdir, *swnirdif, *swddir, *swddni, *swddif;
  float *swupflx, *swupflxc, *swupflxcln, *swdnflx, *swdnflxc, *swdnflxcln;
  float *swuptclnc, *swdntclnc, *swupbclnc, *swdnbclnc;
  int *status;              // device error word (first error wins)
  DebugTaps dbg;
};

struct LwArgs {
  Geo geo;
  DevTables tb;
  CloudFields cf;
  int nlay;                 // LW nlayers
  int ncols;
  const uint32_t *mask;
  int W;
  int variants;
  int o3input, aer_ra_feedback;
  const float *t8w, *p3d, *p8w, *pi3d, *o33d, *tsk, *emiss;
  const float *tauaerlw[16];
  float *rthratenlw, *glw, *olr, *lwcf;
  float *lwupt, *lwuptc, *lwuptcln, *lwdnt, *lwdntc, *lwdntcln, *lwupb, *lwupbc, *lwupbcln, *lwdnb, *lwdnbc, *lwdnbcln;
  float *lwupflx, *lwupflxc, *lwupflxcln, *lwdnflx, *lwdnflxc, *lwdnflxcln;
  float *lwuptclnc, *lwdntclnc, *lwupbclnc, *lwdnbclnc;
  int *status;
  DebugTaps dbg;
};

struct McicaArgs {
  Geo geo;
  int nlay, ngpt, permuteseed, ncols, W, icloud;
  const int *cols;          // nullable: identity
  const float *p3d, *cldfra3d;
  uint32_t *mask;
};

// launchers (defined in the .cu files)
void launch_mcica(const McicaArgs &a, cudaStream_t s);
void launch_sw(const SwArgs &a, int nblocks, cudaStream_t s);
void launch_lw(const LwArgs &a, int nblocks, cudaStream_t s);
void upload_band_descs(const HostTables &T);
int sw_smem_bytes();
int lw_smem_bytes();

}  // namespace arc
