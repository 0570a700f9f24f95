// Kernel argument blocks (device pointers in WRF memory order), per-chunk workspaces and launchers.
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
  __host__ __device__ size_t n2() const { return (size_t)ni * (size_t)(jme - jms + 1); }
  __host__ __device__ size_t n3() const { return (size_t)ni * nk * (size_t)(jme - jms + 1); }
  __host__ __device__ size_t np() const { return (size_t)ni * (nk + 2) * (size_t)(jme - jms + 1); }
  // tile column id -> (i, j)
  __host__ __device__ void ij(int tc, int &i, int &j) const { j = jts + tc / nci; i = its + tc % nci; }
};

struct DebugTaps {     // device pointers (nullable); column index = tile column id
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

// Sweep groups (k_sw_sweep / k_lw_sweep): consecutive g-points of one band handled by one thread per column
constexpr int SWEEP_MAXGRP = 32;
struct SweepGroups { int n; int band[SWEEP_MAXGRP], g0[SWEEP_MAXGRP], ng[SWEEP_MAXGRP]; };
inline SweepGroups make_sweep_groups(const int *ng, const int *g0, int nbands, int gmax) {
  SweepGroups G{};
  for (int b = 0; b < nbands; b++) {
    const int parts = (ng[b] + gmax - 1) / gmax;
    int done = 0;
    for (int q = 0; q < parts; q++) {
      const int n = (ng[b] - done + (parts - q) - 1) / (parts - q);
      G.band[G.n] = b; G.g0[G.n] = g0[b] + done; G.ng[G.n] = n; G.n++;
      done += n;
    }
  }
  return G;
}

constexpr int REC_TILE = 128;                                     // columns per record tile = threads of a sweep block
constexpr int SW_REC = 7 * REC_TILE, SW_REC_R = 4 * REC_TILE, SW_REC_E = 6 * REC_TILE;   // words per SW level record, offsets of R and E

// flux "kinds" in the partial buffer: full up/down, clear up/down, clean up/down, clean-clear up/down
enum { K_FU = 0, K_FD, K_CU, K_CD, K_NU, K_ND, K_XU, K_XD, NKIND };

// Layout of the per-layer coefficient workspace: [layer][column tile of 32][field][32 lanes].  A thread's N fields of one
// layer are N words 128 bytes apart, so the solver addresses them with immediate offsets from one pointer per layer
// (a [field][layer][column] layout costs an address computation per load: 10 % of k_lw_solve's instructions).
// `cap` (column capacity) is a multiple of 256.
__host__ __device__ inline size_t coef_index(int f, int lay, size_t c, size_t cap, int nfields) {
  return (((size_t)lay * (cap >> 5) + (c >> 5)) * nfields + f) * 32 + (c & 31);
}

// ---- SW workspace fields ---------------------------------------------------------------------------
enum { SWC_FAC00 = 0, SWC_FAC01, SWC_FAC10, SWC_FAC11, SWC_H2O, SWC_CO2, SWC_O3, SWC_CH4, SWC_O2, SWC_MOL,
       SWC_SELFFAC, SWC_SELFFRAC, SWC_FORFAC, SWC_FORFRAC, SWC_IDX, SWC_N };
// per-column floats [field][c]
enum { SWF_MU0 = 0, SWF_ALBDIR_NIR, SWF_ALBDIF_NIR, SWF_ALBDIR_UV, SWF_ALBDIF_UV, SWF_ADJFLUX, SWF_N };

struct SwWs {
  int cap;                 // column stride of the per-column / per-layer workspace (outer chunk capacity)
  int pcap;                // column stride of the partial-flux buffers (inner chunk capacity)
  int nlay;                // kte-kts+2
  int W;                   // mask words per (g, column)
  int *cols;               // [cap] tile column id of each chunk column
  float *coef;             // coef_index(field, layer, c, cap, SWC_N)
  float *aer;              // [14][3][nlay][cap]   tau, ssa, asy
  float *cld;              // [14][4][nlay][cap]   taucmc, ssacmc, asmcmc, taormc
  uint32_t *mask;          // [NGSW][W][cap]       McICA bits, bit (lay%32) of word lay/32
  uint32_t *anyc;          // [W][cap]             OR over g of the mask
  int *laytrop;            // [cap]
  int *laysol;             // [14][cap]            layer (0-based) where sfluxzen is taken, -1 = never
  float *colf;             // [SWF_N][cap]
  // Level records handed from k_sw_solve (taumol, reftra, bottom-up sweep) to k_sw_sweep (top-down sweep + band sum),
  // tiled [128-column tile][level][stream slot][g-point][SW_REC words]: one record = 128 lanes x (float4 P | float2 R | float E)
  //   P = (ref, refd, tra, trad) of the layer below the level (level 0 unused), words 0..511
  //   R = (rup, rupd) at the level (level 0 = surface albedos),                 words 512..767
  //   E = direct-beam transmittance of that layer,                              words 768..895
  // (128 columns = one sweep block: its loads are 2 KB / 1 KB / 512 B contiguous pieces)
  // so the solver (one g, all streams) and the sweep (one stream, the g-points of a group) both address a level's records
  // with immediate offsets from one pointer.  Stream slots in the order clear, full [, clean][, clean-clear].
  size_t rec_n;            // words per buffer of the level records (host-side: two buffers are carved)
  float *rec;
  float *zinc;             // [NGSW][pcap]        incident flux of the g-point (adjflux x sfluxzen x mu0)
  float *bpart;            // [sweep group][nlay+1][nk][pcap]  per-group sums of the fluxes; nk = kinds in use, slot of kind k = kslot[k]
  int nk; int kslot[NKIND];
  float *dirs;             // [NGSW][cap]          surface direct beam without delta scaling (x incident flux)
  float *uvni;             // [sweep group][2][pcap] running sums of the FULL surface downward flux over the UV/visible and the near-IR g-points
};

struct SwArgs {
  Geo geo;
  DevTables tb;
  CloudFields cf;
  SwWs ws;
  int ncols;                // columns in this chunk
  int ngroups;              // sweep groups (sw_sweep_groups())
  int variants;             // ARC_VAR_* mask
  int o3input, aer_ra_feedback, sf_surface_physics;
  float solcon;
  const float *t8w, *p3d, *p8w, *pi3d, *o33d, *tsk;
  const float *tauaer300, *tauaer400, *tauaer600, *tauaer999, *gaer400, *gaer600, *waer400, *waer600;
  const float *tauaer3d_sw, *ssaaer3d_sw, *asyaer3d_sw;
  const float *aerod;       // (i,k,j,1:6), non-null = aer_opt 1 (iaer = 6)
  const float *xcoszen, *albedo, *alswvisdir, *alswvisdif, *alswnirdir, *alswnirdif;
  // outputs
  float *rthratensw, *gsw, *swcf, *coszr;
  float *swupt, *swuptc, *swuptcln, *swdnt, *swdntc, *swdntcln, *swupb, *swupbc, *swupbcln, *swdnb, *swdnbc, *swdnbcln;
  float *swvisdir, *swvisdif, *swnirdir, *swnirdif, *swddir, *swddni, *swddif;
  float *swupflx, *swupflxc, *swupflxcln, *swdnflx, *swdnflxc, *swdnflxcln;
  float *swuptclnc, *swdntclnc, *swupbclnc, *swdnbclnc;
  int *status;              // device error word (first error wins)
  DebugTaps dbg;
};

// ---- LW workspace ------------------------------------------------------------------------------------------
enum { LWC_FAC00 = 0, LWC_FAC01, LWC_FAC10, LWC_FAC11, LWC_H2O, LWC_CO2, LWC_O3, LWC_N2O, LWC_CO, LWC_CH4, LWC_O2, LWC_BRD,
       LWC_SELFFAC, LWC_SELFFRAC, LWC_FORFAC, LWC_FORFRAC, LWC_MINORFRAC, LWC_SCALEMINOR, LWC_SCALEMINORN2,
       LWC_PAVEL, LWC_COLDRY, LWC_TAVEL, LWC_TZ, LWC_IDX, LWC_N };
enum { LWF_TZ0 = 0, LWF_TBOUND, LWF_EMISS, LWF_N };

struct LwWs {
  int cap, pcap, nlay, W;
  int *cols;               // [cap] tile column id of each chunk column (cloud-bucketed list); nullable = identity from col0
  float *coef;             // coef_index(field, layer, c, cap, LWC_N)
  float *aer;              // [16][nlay][cap]
  float *cld;              // [16][nlay][cap]  taucmc
  uint32_t *mask;          // [NGLW][W][cap]
  uint32_t *anyc;          // [W][cap]
  int *laytrop;            // [cap]
  float *colf;             // [LWF_N][cap]
  float *secdiff;          // [16][cap]
  // Records k_lw_band's downward pass leaves for its upward pass, [layer][g-point][pcap] float4 (a warp = 512 contiguous bytes):
  //   rec  (atrans, bbu) of the full stream, (atrans, bbu) of the clean stream: radlu' = radlu + (bbu - radlu) atrans
  //   recC (X, Y) x (full, clean) of radlu' = radlu - radlu X + Y; written and read only where the column has cloud in the layer
  // One buffer: the thread that wrote a record is the one that reads it.
  float4 *rec, *recC;
  float *bpart;            // [band group][nlay+1][nk][pcap]  sums of the radiances over a group's g-points (k_lw_band -> k_lw_reduce); nk = kinds in use, slot of kind k = kslot[k]
  int nk; int kslot[NKIND];
};

struct LwArgs {
  Geo geo;
  DevTables tb;
  CloudFields cf;
  LwWs ws;
  int col0;                 // first tile column of this chunk
  int ncols;
  int ngroups;              // sweep groups (lw_sweep_groups())
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
  int nlay, nz, ngpt, permuteseed, ncols, W, icloud, cap, col0;
  int lw_buffer;            // 1: LW layering above the model top (4-hPa buffer layers), 0: SW single extra layer
  const int *cols;          // nullable: tile column = col0 + c
  const float *p3d, *p8w, *cldfra3d;
  uint32_t *mask;           // [ngpt][W][cap]
  uint32_t *anyc;           // [W][cap]
};

// launchers (defined in the .cu files); every launcher bumps the launch counter
void launch_compact_sunlit(const Geo &g, const float *xcoszen, const float *cldfra3d, int *cols, int *count, int slot, cudaStream_t s);
void launch_sw_night(const SwArgs &a, cudaStream_t s);
void launch_mcica(const McicaArgs &a, cudaStream_t s);
void launch_sw_prep(const SwArgs &a, cudaStream_t s);
void launch_sw_solve(const SwArgs &a, cudaStream_t s);
void launch_sw_sweep(const SwArgs &a, cudaStream_t s);
int sw_sweep_groups();
int lw_sweep_groups();
void launch_sw_reduce(const SwArgs &a, cudaStream_t s);
void launch_lw_prep(const LwArgs &a, cudaStream_t s);
void launch_cal_cldfra1(const Geo &G, const float *qv, const float *qc, const float *qi, const float *qs, int f_qv, int f_qc, int f_qi, int f_qs,
                        const float *t_phy, const float *p_phy, const float *f_ice_phy, int mp_physics, float *cldfra, int *flag, cudaStream_t s);
constexpr int ARC_OZN_MAXLEV = 128;      // data levels of the ozone climatology passed by value to the kernel (CAM: 59)
void launch_cal_cldfra3(const Geo &G, float *cldfra, const float *qv, float *qc, float *qi, const float *qs, const float *p, const float *t,
                        const float *rho, const float *xland, float gridkm, float *qvsat, float *theta, float *dz, cudaStream_t s);
void launch_cal_cldfra2(const Geo &G, const float *qc, const float *qi, int f_qc, int f_qi, float *cldfra, cudaStream_t s);
void launch_ozn_time_int(const Geo &G, int levsiz, const float *m0, const float *m1, float fact1, float fact2, float *ozmixt, cudaStream_t s);
void launch_clim_p_int(const Geo &G, int levsiz, const float *pin_host, const float *p, float pscale, int nsrc, const float *data, float *out,
                       const float *pf, float *total, cudaStream_t s);
void launch_ozn_p_int(const Geo &G, int levsiz, const float *pin_host, const float *p, const float *ozmixt, float *o3vmr, cudaStream_t s);
void launch_lw_band(const LwArgs &a, cudaStream_t s);
bool lw_layout_ok();
void launch_lw_reduce(const LwArgs &a, cudaStream_t s);
void upload_band_descs(const HostTables &T);
void launch_selftest_pt(const DevTables &tb, const float *p, const float *t, int n, int *packed, cudaStream_t s);
void count_launch(int n = 1);
long long launch_count();

}  // namespace arc
