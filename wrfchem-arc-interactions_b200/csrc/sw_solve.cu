// Shortwave spectral solver: k_sw_solve = taumol_sw + layer optics + reftra_sw + bottom-up half of vrtqdr_sw for all call
// variants, k_sw_sweep = top-down half + fluxes + ordered sum over the g-points of a band,
// k_sw_reduce = band sum / heating rates / scatter.
//
// Reference (module_ra_rrtmg_sw.F v3.9.1): taumol_sw 3081-4540, reftra_sw 2422-2701, vrtqdr_sw 7922-8046,
// spcvmc_sw 8083-8658, rrtmg_sw 9376-9478 (full+clear call, then the clean call with ztauacln = 0),
// RRTMG_SWRAD output scatter 11125-11172.
//
// Variants ("streams") and the layer evaluations they share.  McICA masks are binary, so in every layer a
// stream uses either the clear-layer or the cloudy-layer two-stream solution:
//     stream       layer not cloudy        layer cloudy
//     CLEAR        Pa  (gas+ray+aerosol)   Pa
//     FULL         Pa                      Pca (Pa optics + cloud)
//     CLEAN        Pn  (gas+ray)           Pcn (Pn optics + cloud)
//     CLEANCLEAR   Pn                      Pn
// so one layer costs 2 reftra evaluations (3 or 4 when cloudy) instead of the reference's 2 x 2.
#include "args.h"
#include "../../include/arc_rad.h"

namespace arc {


static __constant__ SwBandDesc c_sw[14];
static __constant__ int c_sw_ngb[NGSW];    // band index 0..13 of each SW g-point
// Sweep groups: the g-points of a band are swept by one thread per column in groups of at most SWEEP_GMAX consecutive
// g-points (register state); bands with more g-points are cut into two equal groups.
#ifndef SW_GMAX
#define SW_GMAX 12
#endif
static SweepGroups h_sw_grp;
static __constant__ int c_sw_grp_band[SWEEP_MAXGRP];
void upload_band_descs_sw(const HostTables &T) {
  cudaMemcpyToSymbol(c_sw, T.sw, sizeof(SwBandDesc) * 14);
  int ngs[14], g0s[14];
  for (int b = 0; b < 14; b++) { ngs[b] = T.sw[b].ng; g0s[b] = T.sw[b].g0; }
  h_sw_grp = make_sweep_groups(ngs, g0s, 14, SW_GMAX);
  cudaMemcpyToSymbol(c_sw_grp_band, h_sw_grp.band, sizeof(int) * SWEEP_MAXGRP);
  int ngb[NGSW];
  for (int i = 0; i < NGSW; i++) ngb[i] = T.sw_ngb[i] - 1;
  cudaMemcpyToSymbol(c_sw_ngb, ngb, sizeof(int) * NGSW);
}

struct SwLay {
  float fac00, fac01, fac10, fac11, h2o, co2, o3, ch4, o2, mol, selffac, selffrac, forfac, forfrac;
  int jp, jt, jt1, indself, indfor;
};

struct SwBin { float speccomb, fs; int js, ind0, ind1; };

__device__ __forceinline__ SwBin sw_binary(const SwLay &L, float cola, float colb_scaled, float mult, int nsp, bool lower,
                                           float oneminus) {
  SwBin b;
  b.speccomb = __fadd_rn(cola, colb_scaled);
  float specparm = div_rn(cola, b.speccomb);
  if (specparm >= oneminus) specparm = oneminus;
  const float specmult = __fmul_rn(mult, specparm);
  b.js = 1 + (int)specmult;
  b.fs = fmod1(specmult);
  if (lower) {
    b.ind0 = ((L.jp - 1) * 5 + (L.jt - 1)) * nsp + b.js;
    b.ind1 = (L.jp * 5 + (L.jt1 - 1)) * nsp + b.js;
  } else {
    b.ind0 = ((L.jp - 13) * 5 + (L.jt - 1)) * nsp + b.js;
    b.ind1 = ((L.jp - 12) * 5 + (L.jt1 - 1)) * nsp + b.js;
  }
  return b;
}

// 8-point (eta, T, p) interpolation; `ab` points at element (1) of the g-point's column, st = 9 (lower) / 5 (upper)
__device__ __forceinline__ float sw_k8(const float *__restrict__ ab, const SwLay &L, const SwBin &b, int st) {
  const float f1 = 1.f - b.fs;
  const float *p0 = ab + b.ind0 - 1, *p1 = ab + b.ind1 - 1;
  return (f1 * L.fac00) * p0[0] + (b.fs * L.fac00) * p0[1] + (f1 * L.fac10) * p0[st] + (b.fs * L.fac10) * p0[st + 1] +
         (f1 * L.fac01) * p1[0] + (b.fs * L.fac01) * p1[1] + (f1 * L.fac11) * p1[st] + (b.fs * L.fac11) * p1[st + 1];
}
__device__ __forceinline__ float sw_k4(const float *__restrict__ ab, const SwLay &L, bool lower) {
  int ind0, ind1;
  if (lower) { ind0 = ((L.jp - 1) * 5 + (L.jt - 1)); ind1 = (L.jp * 5 + (L.jt1 - 1)); }
  else { ind0 = ((L.jp - 13) * 5 + (L.jt - 1)); ind1 = ((L.jp - 12) * 5 + (L.jt1 - 1)); }
  return L.fac00 * ab[ind0] + L.fac10 * ab[ind0 + 1] + L.fac01 * ab[ind1] + L.fac11 * ab[ind1 + 1];
}

// Gas optical depth, Rayleigh optical depth and (when this is the band's source layer) the solar source of one
// (layer, g-point).  S = the g-point's table slice in shared memory, D = band descriptor.
__device__ __forceinline__ void sw_taumol(int band, const float *__restrict__ S, const SwBandDesc &D, const SwLay &L, bool lower,
                                          float oneminus, float &taug, float &taur, float &sflux) {
  const float *A = S + D.oA, *B = S + D.oB;
  auto selfk = [&]() { const float *r = S + D.oSelf + L.indself - 1; return L.selffac * (r[0] + L.selffrac * (r[1] - r[0])); };
  auto fork = [&]() { const float *r = S + D.oFor + L.indfor - 1; return L.forfac * (r[0] + L.forfrac * (r[1] - r[0])); };
  auto sflx_eta = [&](const SwBin &b) { const float *r = S + D.oSflx + b.js - 1; return r[0] + b.fs * (r[1] - r[0]); };
  const float rayl0 = S[D.oRayl];
  taur = L.mol * rayl0;
  switch (band) {
    case 16:
      if (lower) { SwBin b = sw_binary(L, L.h2o, __fmul_rn(D.strrat, L.ch4), 8.f, 9, true, oneminus);
                   taug = b.speccomb * sw_k8(A, L, b, 9) + L.h2o * (selfk() + fork()); }
      else { taug = L.ch4 * sw_k4(B, L, false); sflux = S[D.oSflx]; }
      break;
    case 17:
      if (lower) { SwBin b = sw_binary(L, L.h2o, __fmul_rn(D.strrat, L.co2), 8.f, 9, true, oneminus);
                   taug = b.speccomb * sw_k8(A, L, b, 9) + L.h2o * (selfk() + fork()); }
      else { SwBin b = sw_binary(L, L.h2o, __fmul_rn(D.strrat, L.co2), 4.f, 5, false, oneminus);
             taug = b.speccomb * sw_k8(B, L, b, 5) + L.h2o * fork(); sflux = sflx_eta(b); }
      break;
    case 18: case 19: {
      const float colb = band == 18 ? L.ch4 : L.co2;
      if (lower) { SwBin b = sw_binary(L, L.h2o, __fmul_rn(D.strrat, colb), 8.f, 9, true, oneminus);
                   taug = b.speccomb * sw_k8(A, L, b, 9) + L.h2o * (selfk() + fork()); sflux = sflx_eta(b); }
      else taug = colb * sw_k4(B, L, false);
      break; }
    case 20:
      if (lower) { taug = L.h2o * (sw_k4(A, L, true) + selfk() + fork()) + L.ch4 * S[D.oMisc]; sflux = S[D.oSflx]; }
      else taug = L.h2o * (sw_k4(B, L, false) + fork()) + L.ch4 * S[D.oMisc];
      break;
    case 21:
      if (lower) { SwBin b = sw_binary(L, L.h2o, __fmul_rn(D.strrat, L.co2), 8.f, 9, true, oneminus);
                   taug = b.speccomb * sw_k8(A, L, b, 9) + L.h2o * (selfk() + fork()); sflux = sflx_eta(b); }
      else { SwBin b = sw_binary(L, L.h2o, __fmul_rn(D.strrat, L.co2), 4.f, 5, false, oneminus);
             taug = b.speccomb * sw_k8(B, L, b, 5) + L.h2o * fork(); }
      break;
    case 22: {
      const float o2adj = 1.6f;
      const float o2cont = 4.35e-4f * L.o2 / (350.0f * 2.0f);
      if (lower) { SwBin b = sw_binary(L, L.h2o, __fmul_rn(__fmul_rn(o2adj, D.strrat), L.o2), 8.f, 9, true, oneminus);
                   taug = b.speccomb * sw_k8(A, L, b, 9) + L.h2o * (selfk() + fork()) + o2cont; sflux = sflx_eta(b); }
      else taug = L.o2 * o2adj * sw_k4(B, L, false) + o2cont;
      break; }
    case 23:
      if (lower) { taug = L.h2o * (D.givfac * sw_k4(A, L, true) + selfk() + fork()); sflux = S[D.oSflx]; }
      else taug = 0.f;
      break;
    case 24:
      if (lower) { SwBin b = sw_binary(L, L.h2o, __fmul_rn(D.strrat, L.o2), 8.f, 9, true, oneminus);
                   const float *r = S + D.oRayl + b.js - 1;
                   taur = L.mol * (r[0] + b.fs * (r[1] - r[0]));
                   taug = b.speccomb * sw_k8(A, L, b, 9) + L.o3 * S[D.oMisc] + L.h2o * (selfk() + fork()); sflux = sflx_eta(b); }
      else { taur = L.mol * S[D.oRayl + 9]; taug = L.o2 * sw_k4(B, L, false) + L.o3 * S[D.oMisc + 1]; }
      break;
    case 25:
      if (lower) { taug = L.h2o * sw_k4(A, L, true) + L.o3 * S[D.oMisc]; sflux = S[D.oSflx]; }
      else taug = L.o3 * S[D.oMisc + 1];
      break;
    case 26:
      taug = 0.f; sflux = S[D.oSflx];
      break;
    case 27:
      if (lower) taug = L.o3 * sw_k4(A, L, true);
      else { taug = L.o3 * sw_k4(B, L, false); sflux = D.scalekur * S[D.oSflx]; }
      break;
    case 28:
      if (lower) { SwBin b = sw_binary(L, L.o3, __fmul_rn(D.strrat, L.o2), 8.f, 9, true, oneminus); taug = b.speccomb * sw_k8(A, L, b, 9); }
      else { SwBin b = sw_binary(L, L.o3, __fmul_rn(D.strrat, L.o2), 4.f, 5, false, oneminus);
             taug = b.speccomb * sw_k8(B, L, b, 5); sflux = sflx_eta(b); }
      break;
    default:  // 29
      if (lower) taug = L.h2o * (sw_k4(A, L, true) + selfk() + fork()) + L.co2 * S[D.oMisc];
      else { taug = L.co2 * sw_k4(B, L, false) + L.h2o * S[D.oMisc + 1]; sflux = S[D.oSflx]; }
      break;
  }
}

// exp(-x) through the reference's Pade-indexed table (SW:2590-2600, 8445-8460): series below od_lo
__device__ __forceinline__ float sw_expt(const float *__restrict__ exp_tbl, float x, float bpade) {
  // both branches are evaluated and selected (the table index of a small x is valid): no divergence, more ILP
  const float ser = __fadd_rn(__fsub_rn(1.f, x), __fmul_rn(__fmul_rn(0.5f, x), x));
  const float tblind = div_rn(x, __fadd_rn(bpade, x));
  const int itind = (int)__fadd_rn(__fmul_rn(10000.0f, tblind), 0.5f);
  const float tab = exp_tbl[itind];
  return x <= 0.06f ? ser : tab;
}

// Unfused IEEE single-precision operations.  reftra_sw and the optical-property mixing that feeds it are evaluated with
// exactly the reference's operation order and rounding: for nearly conservative layers (1 - w ~ 1e-6, common in the
// aerosol-free "clean" stream of the Rayleigh-dominated bands) k = sqrt(g1^2 - g2^2) and the numerators of R and T are
// differences of nearly equal O(1) numbers, so a contracted FMA or an approximate reciprocal would move the layer
// reflectance by percents relative to the reference.  The adding recurrences (vrtqdr_sw) are well conditioned and keep
// the fast reciprocal.
#define M_(a, b) __fmul_rn((a), (b))
#define A_(a, b) __fadd_rn((a), (b))
#define S_(a, b) __fsub_rn((a), (b))
#define D_(a, b) div_rn((a), (b))
#define R_(b) rcp_rn(b)

// reftra_sw (kmodts = 2, PIFM) for one layer, SW:2540-2690.  Returns (ref, refd, tra, trad) and, because both branches
// evaluate it anyway, e = exp(-tau/mu0) through the table: the direct-beam transmittance spcvmc_sw computes again from the
// same operands (SW:8560-8575), identical bit for bit as long as tau/mu0 <= 500 (reftra clamps its argument there).
// ZG0: the asymmetry parameter is exactly zero (Rayleigh + gas layer of the aerosol-free streams); the expressions in zg
// then reduce exactly (0 * x, x - 0, 0 / x, x / 1) and two divisions drop out.
struct SwLayerRT { float4 p; float e; };
#ifndef SW_REFTRA_INLINE
#define SW_REFTRA_INLINE __forceinline__
#endif
template <bool ZG0>
__device__ SW_REFTRA_INLINE SwLayerRT sw_reftra(const float *__restrict__ exp_tbl, float bpade, float zg, float prmuz, float rmuz, float zto1, float zw) {
  const float eps = 1.e-08f, zwcrit = 0.9999995f;
  float4 o;
  const float zx = div_rn_r(zto1, prmuz, rmuz);        // rmuz = rcp_rn(prmuz), hoisted out of the layer loop
  const float zexp = sw_expt(exp_tbl, fminf(zx, 500.f), bpade);
  float zgamma1, zgamma2, zgamma3, zgamma4, zwo;
  if (ZG0) {
    zgamma1 = M_(S_(8.f, M_(zw, 5.f)), 0.25f);
    zgamma2 = M_(M_(3.f, zw), 0.25f);
    zgamma3 = 0.5f; zgamma4 = 0.5f;
    zwo = zw;
  } else {
    const float zg3 = M_(3.f, zg);
    zgamma1 = M_(S_(8.f, M_(zw, A_(5.f, zg3))), 0.25f);
    zgamma2 = M_(M_(3.f, M_(zw, S_(1.f, zg))), 0.25f);
    zgamma3 = M_(S_(2.f, M_(zg3, prmuz)), 0.25f);
    zgamma4 = S_(1.f, zgamma3);
    const float q = D_(zg, S_(1.f, zg));
    const float denom = fmaxf(S_(1.f, M_(S_(1.f, zw), M_(q, q))), 1.0E-30f);
    zwo = D_(zw, denom);
  }
  if (zwo >= zwcrit) {
    const float za = M_(zgamma1, prmuz);
    const float za1 = S_(za, zgamma3);
    const float zgt = M_(zgamma1, zto1);
    const float ze2 = zexp;
    o.x = D_(S_(zgt, M_(za1, S_(1.f, ze2))), A_(1.f, zgt));
    o.z = S_(1.f, o.x);
    o.y = D_(zgt, A_(1.f, zgt));
    o.w = S_(1.f, o.y);
    if (ze2 == 1.0f) { o.x = 0.f; o.z = 1.f; o.y = 0.f; o.w = 1.f; }
  } else {
    const float za1 = A_(M_(zgamma1, zgamma4), M_(zgamma2, zgamma3));
    const float za2 = A_(M_(zgamma1, zgamma3), M_(zgamma2, zgamma4));
    const float zrk = __fsqrt_rn(S_(M_(zgamma1, zgamma1), M_(zgamma2, zgamma2)));
    const float zrp = M_(zrk, prmuz);
    const float zrp1 = A_(1.f, zrp), zrm1 = S_(1.f, zrp), zrk2 = M_(2.f, zrk);
    const float zrpp = S_(1.f, M_(zrp, zrp));
    const float zrkg = A_(zrk, zgamma1);
    const float zr1 = M_(zrm1, A_(za2, M_(zrk, zgamma3)));
    const float zr2 = M_(zrp1, S_(za2, M_(zrk, zgamma3)));
    const float zr3 = M_(zrk2, S_(zgamma3, M_(za2, prmuz)));
    const float zr4 = M_(zrpp, zrkg);
    const float zr5 = M_(zrpp, S_(zrk, zgamma1));
    const float zt1 = M_(zrp1, A_(za1, M_(zrk, zgamma4)));
    const float zt2 = M_(zrm1, S_(za1, M_(zrk, zgamma4)));
    const float zt3 = M_(zrk2, A_(zgamma4, M_(za1, prmuz)));
    const float zbeta = D_(S_(zgamma1, zrk), zrkg);
    const float ze1 = fminf(M_(zrk, zto1), 500.f);
    const float zem1 = sw_expt(exp_tbl, ze1, bpade), zep1 = R_(zem1);
    const float zem2 = zexp, zep2 = R_(zem2);
    const float zdenr = A_(M_(zr4, zep1), M_(zr5, zem1));
    const float zdent = zdenr;   // zt4 = zr4, zt5 = zr5
    if (zdenr >= -eps && zdenr <= eps) { o.x = eps; o.z = zem2; }
    else {
      o.x = D_(M_(zw, S_(S_(M_(zr1, zep1), M_(zr2, zem1)), M_(zr3, zem2))), zdenr);
      o.z = S_(zem2, D_(M_(M_(zem2, zw), S_(S_(M_(zt1, zep1), M_(zt2, zem1)), M_(zt3, zep2))), zdent));
    }
    const float zemm = M_(zem1, zem1);
    const float zdend = R_(M_(S_(1.f, M_(zbeta, zemm)), zrkg));
    o.y = M_(M_(zgamma2, S_(1.f, zemm)), zdend);
    o.w = M_(M_(zrk2, zem1), zdend);
  }
  SwLayerRT r;
  r.p = o;
  r.e = zx <= 500.f ? zexp : sw_expt(exp_tbl, zx, bpade);      // beyond the clamp the table is 0 or expeps
  return r;
}

// ------------------------------------------------------------------------------------------------------
#ifndef SW_MINBLOCKS
#define SW_MINBLOCKS 4
#endif
template <int NL>
__global__ void __launch_bounds__(256, SW_MINBLOCKS) k_sw_solve(SwArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *s_exp = reinterpret_cast<float *>(smem_raw);                 // 10004 floats
  float *S = s_exp + 10004;                                           // slice
  uint64_t *bar = reinterpret_cast<uint64_t *>(S + SLICE_MAX);

  // Block order: band-major, then column tile, then g-point within the band.  Blocks resident at the same time run the
  // same band (one code path in the instruction cache) and the g-points of a band share the tile's workspace lines in L2.
  const int ntiles = (a.ncols + 255) >> 8;
  int b = 0;
  while (b < NBSW - 1 && (int)blockIdx.x >= c_sw[b + 1].g0 * ntiles) b++;
  const SwBandDesc &D = c_sw[b];
  const int rblk = blockIdx.x - D.g0 * ntiles;
  const int tile = rblk / D.ng;
  const int g = D.g0 + rblk % D.ng;
  const int band = b + 16;
  {
    StageReq req[2] = {{s_exp, a.tb.sw_exp, 10004 * 4},
                       {S, a.tb.sw_tab + D.slice_base + (size_t)D.slice_floats * (g - D.g0), (uint32_t)D.slice_floats * 4}};
    stage_tables(bar, req, 2);
  }
  const int c = tile * 256 + threadIdx.x;
  if (c >= a.ncols) return;

  const SwWs &ws = a.ws;
  const int nlay = ws.nlay;
  const size_t cap = ws.cap;
  const float bpade = a.tb.bpade, oneminus = a.tb.oneminus;
  const bool do_clean = (a.variants & ARC_VAR_CLEAN) != 0;
  const bool do_clnc = (a.variants & ARC_VAR_CLEANCLEAR) != 0;
  const bool noaer = do_clean || do_clnc;

  const float prmu0 = ws.colf[(size_t)SWF_MU0 * cap + c];
  const float rmu0 = rcp_rn(prmu0);
  const bool uv = (b >= 9 && b <= 12);
  const float albp = ws.colf[(size_t)(uv ? SWF_ALBDIR_UV : SWF_ALBDIR_NIR) * cap + c];
  const float albd = ws.colf[(size_t)(uv ? SWF_ALBDIF_UV : SWF_ALBDIF_NIR) * cap + c];
  const int laytrop = ws.laytrop[c];
  const int laysol = ws.laysol[(size_t)b * cap + c];

  uint32_t mw[NL / 32];
#pragma unroll
  for (int w = 0; w < NL / 32; w++) mw[w] = w < ws.W ? ws.mask[((size_t)g * ws.W + w) * cap + c] : 0u;

  // Level records handed to k_sw_sweep (layout in args.h): for every stream the two-stream solution (P, e) of the layer
  // below the level and the upward reflectances (rup, rupd) at the level.  Stream order: clear, full, clean, clean-clear.
  const size_t pcap = ws.pcap;
  const int slot[4] = {0, 1, 2, do_clean ? 3 : 2};                     // compact stream slots
  const int nstream = 2 + (do_clean ? 1 : 0) + (do_clnc ? 1 : 0);
  const unsigned lvstride = (unsigned)nstream * NGSW * SW_REC;         // words per (tile, level)
  // record of (this tile, level 0, stream slot 0, g), lane part added per field
  float *rec = ws.rec + ((size_t)(c / REC_TILE) * (nlay + 1) * nstream * NGSW + g) * SW_REC;
  const int lane = c % REC_TILE;
  float rup[4], rupd[4];
#pragma unroll
  for (int s = 0; s < 4; s++) {
    rup[s] = albp; rupd[s] = albd;
    if (s == 2 && !do_clean) continue;
    if (s == 3 && !do_clnc) continue;
    reinterpret_cast<float2 *>(rec + slot[s] * (NGSW * SW_REC) + SW_REC_R)[lane] = make_float2(albp, albd);
  }
  float sfluxzen = 0.f;
  // un-delta-scaled direct transmittance of every layer of the FULL stream (thread-private, 4 B per layer): the reference
  // multiplies them from the top down (ztdbt_nodel, SW:8560-8575) while this loop runs bottom-up, and the product is rounded
  // at every step - so they are kept and multiplied in the reference's order after the loop
  float enod[NL];

  // all workspace words of a layer are requested together at the top of the iteration (one wait per layer)
  const float *coef = ws.coef + coef_index(0, 0, c, cap, SWC_N);
  const unsigned lstride = (unsigned)cap * SWC_N;                                  // coefficient words per layer
  const unsigned stf = (unsigned)nlay * (unsigned)cap, ucap = (unsigned)cap;      // 32-bit offsets: SWC_N*nlay*cap < 2^31
  auto load_layer = [&](int lay, SwLay &L, int &pk, float &ta, float &om, float &as) {
    const float *p = coef + (size_t)((unsigned)lay * lstride);       // fields at immediate offsets of 128 bytes
    L.fac00 = p[SWC_FAC00 * 32]; L.fac01 = p[SWC_FAC01 * 32]; L.fac10 = p[SWC_FAC10 * 32]; L.fac11 = p[SWC_FAC11 * 32];
    L.h2o = p[SWC_H2O * 32]; L.co2 = p[SWC_CO2 * 32]; L.o3 = p[SWC_O3 * 32]; L.ch4 = p[SWC_CH4 * 32]; L.o2 = p[SWC_O2 * 32];
    L.mol = p[SWC_MOL * 32];
    L.selffac = p[SWC_SELFFAC * 32]; L.selffrac = p[SWC_SELFFRAC * 32]; L.forfac = p[SWC_FORFAC * 32]; L.forfrac = p[SWC_FORFRAC * 32];
    pk = __float_as_int(p[SWC_IDX * 32]);
    const float *pa = ws.aer + c + (unsigned)(b * 3) * stf + (unsigned)lay * ucap;
    ta = pa[0]; om = pa[stf]; as = pa[2u * stf];
  };
  for (int lay = 0; lay < nlay; lay++) {
    SwLay L; int pk; float taua, omga, asya;
    load_layer(lay, L, pk, taua, omga, asya);
    L.jp = IDX_JP(pk); L.jt = IDX_JT(pk); L.jt1 = IDX_JT1(pk); L.indself = IDX_SELF(pk); L.indfor = IDX_FOR(pk);
    const bool lower = lay < laytrop;
    float taug, taur, sfl = 0.f;
    sw_taumol(band, S, D, L, lower, oneminus, taug, taur, sfl);
    if (lay == laysol) sfluxzen = sfl;
    if (a.dbg.taug) {
      const size_t q = ((size_t)ws.cols[c] * nlay + lay) * NGSW + g;
      a.dbg.taug[q] = taug; a.dbg.taur[q] = taur;
    }
    const bool cloudy = (mw[lay >> 5] >> (lay & 31)) & 1u;
    float taucmc = 0.f, ssacmc = 1.f, asmcmc = 0.f, taormc = 0.f;
    if (cloudy) {
      const float *pc = ws.cld + c + (unsigned)(b * 4) * stf + (unsigned)lay * ucap;
      taucmc = pc[0]; ssacmc = pc[stf]; asmcmc = pc[2u * stf]; taormc = pc[3u * stf];
    }
    if (a.dbg.taucmc) a.dbg.taucmc[((size_t)ws.cols[c] * nlay + lay) * NGSW + g] = taucmc;

    // ---- optical properties of the clear and cloudy layer, with (v = 0) and without (v = 1) aerosol
    float4 pclr[2], pcld[2];
    float eclr[2], ecld[2];
#pragma unroll
    for (int v = 0; v < 2; v++) {
      if (v == 1 && !noaer) break;
      float ztauc, zomcc, zgcc;
      if (v == 0) {
        ztauc = A_(A_(taur, taug), taua);
        zomcc = A_(M_(taur, 1.0f), M_(taua, omga));
        zgcc = D_(M_(M_(asya, omga), taua), zomcc);
        zomcc = D_(zomcc, ztauc);
        // direct beam without delta scaling (diagnostic surface direct flux of the FULL stream)
        const float tauorig = cloudy ? A_(ztauc, taormc) : ztauc;
        enod[lay] = sw_expt(s_exp, div_rn_r(tauorig, prmu0, rmu0), bpade);
        const float zf = M_(zgcc, zgcc);
        const float zwf = M_(zomcc, zf);
        ztauc = M_(S_(1.0f, zwf), ztauc);
        zomcc = D_(S_(zomcc, zwf), fmaxf(S_(1.0f, zwf), 1.0E-30f));
        zgcc = D_(S_(zgcc, zf), fmaxf(S_(1.0f, zf), 1.0E-30f));
      } else {
        // zero aerosol: the same expressions with taua = 0 reduce exactly (x + 0, x * 1, 0 / x, x / 1) to
        ztauc = A_(taur, taug);
        zomcc = D_(taur, ztauc);
        zgcc = 0.f;
      }
      const SwLayerRT rt = v == 0 ? sw_reftra<false>(s_exp, bpade, zgcc, prmu0, rmu0, ztauc, zomcc)
                                  : sw_reftra<true>(s_exp, bpade, 0.f, prmu0, rmu0, ztauc, zomcc);
      pclr[v] = rt.p; eclr[v] = rt.e;
      if (cloudy) {
        const float ztauo = A_(ztauc, taucmc);
        float zomco = A_(M_(ztauc, zomcc), M_(taucmc, ssacmc));
        const float zgco = D_(A_(M_(M_(taucmc, ssacmc), asmcmc), M_(M_(ztauc, zomcc), zgcc)), zomco);
        zomco = D_(zomco, ztauo);
        const SwLayerRT rc = sw_reftra<false>(s_exp, bpade, zgco, prmu0, rmu0, ztauo, zomco);
        pcld[v] = rc.p; ecld[v] = rc.e;
      }
    }
    // ---- upward reflectances at the top of this layer (vrtqdr_sw bottom-up sweep)
#pragma unroll
    for (int s = 0; s < 4; s++) {
      if (s == 2 && !do_clean) continue;
      if (s == 3 && !do_clnc) continue;
      float4 P; float e;
      if (s == 0) { P = pclr[0]; e = eclr[0]; }
      else if (s == 1) { P = cloudy ? pcld[0] : pclr[0]; e = cloudy ? ecld[0] : eclr[0]; }
      else if (s == 2) { P = cloudy ? pcld[1] : pclr[1]; e = cloudy ? ecld[1] : eclr[1]; }
      else { P = pclr[1]; e = eclr[1]; }
      // vrtqdr_sw's bottom-up recurrence (SW:7997-8005) in the reference's own operation order and rounding (unfused, IEEE
      // reciprocal): with reftra_sw's operands bit-exact the whole shortwave chain reproduces the reference bit for bit,
      // and streams with identical inputs (zero aerosol: clean == full; no cloud: clear == full) stay identical
      const float zreflect = R_(S_(1.f, M_(rupd[s], P.y)));
      const float nrup = A_(P.x, M_(M_(P.w, A_(M_(S_(P.z, e), rupd[s]), M_(e, rup[s]))), zreflect));
      const float nrupd = A_(P.y, M_(M_(M_(P.w, P.w), rupd[s]), zreflect));
      rup[s] = nrup; rupd[s] = nrupd;
      float *q = rec + (size_t)((unsigned)(lay + 1) * lvstride) + slot[s] * (NGSW * SW_REC);
      // (where its layer is not cloudy the FULL stream's two-stream solution IS the CLEAR one: k_sw_sweep reads it there)
      if (s != 1 || cloudy) { reinterpret_cast<float4 *>(q)[lane] = P; q[SW_REC_E + lane] = e; }
      reinterpret_cast<float2 *>(q + SW_REC_R)[lane] = make_float2(nrup, nrupd);
    }
  }
  if (a.dbg.sfluxzen) a.dbg.sfluxzen[(size_t)ws.cols[c] * NGSW + g] = sfluxzen;

  // incident flux of this g-point and the un-delta-scaled direct beam at the surface; the top-down sweep runs in k_sw_sweep
  float tdir_nodel = 1.f;
  for (int lay = nlay - 1; lay >= 0; lay--) tdir_nodel = M_(enod[lay], tdir_nodel);
  const float zincflx = ws.colf[(size_t)SWF_ADJFLUX * cap + c] * sfluxzen * prmu0;
  ws.zinc[(size_t)g * pcap + c] = zincflx;
  ws.dirs[(size_t)g * pcap + c] = __fmul_rn(zincflx, tdir_nodel);
}

static int sw_solve_smem() { return (10004 + SLICE_MAX) * 4 + 16; }

void launch_sw_solve(const SwArgs &a, cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_sw_solve<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, sw_solve_smem());
    cudaFuncSetAttribute(k_sw_solve<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, sw_solve_smem());
    cudaFuncSetAttribute(k_sw_solve<160>, cudaFuncAttributeMaxDynamicSharedMemorySize, sw_solve_smem());
    attr = true;
  }
  dim3 grid(NGSW * ((a.ncols + 255) / 256));
  if (a.ws.nlay <= 64) k_sw_solve<64><<<grid, 256, sw_solve_smem(), s>>>(a);
  else if (a.ws.nlay <= 128) k_sw_solve<128><<<grid, 256, sw_solve_smem(), s>>>(a);
  else k_sw_solve<160><<<grid, 256, sw_solve_smem(), s>>>(a);
  count_launch();
}

// ------------------------------------------------------------------------------------------------------
// Top-down sweep of vrtqdr_sw (SW:8007-8045) + ordered sum over the g-points of a band (SW:8617-8650).
// One thread per (column, band, stream): the (tdbt, tdn, rdnd) of the band's NG g-points are its register state; per level
// it requests the 3 NG records of that level together (written by k_sw_solve; the addresses do not depend on the
// recurrence), forms the NG up / down fluxes, adds them in g order and writes ONE band partial [band][level][kind][c] per
// kind.  No shared memory, no barriers, no atomics; lanes = neighbouring columns.  HBM-bound: 28 B per (column, g, level,
// stream).
template <int NG>
__global__ void __launch_bounds__(128) k_sw_sweep(SwArgs a, int grp, int g0, int nstream) {
  const SwWs &ws = a.ws;
  // blocks of the streams of one tile are neighbours in launch order: the FULL block re-reads the CLEAR records of its
  // non-cloudy layers while they are still in L2
  const int tile = blockIdx.x / nstream;
  const int si = blockIdx.x % nstream;         // compact stream slot
  const int c = tile * 128 + threadIdx.x;
  if (c >= a.ncols) return;
  const bool do_clean = (a.variants & ARC_VAR_CLEAN) != 0;
  const int s = si < 2 ? si : (si == 2 && do_clean ? 2 : 3);      // 0 clear, 1 full, 2 clean, 3 clean-clear
  const int nlay = ws.nlay, nk = ws.nk;
  const size_t pcap = ws.pcap;
  const int lane = threadIdx.x;                                        // block = one record tile
  const unsigned lvstride = (unsigned)nstream * NGSW * SW_REC;         // words per (tile, level)
  const float *__restrict__ rec = ws.rec + (((size_t)tile * (nlay + 1) * nstream + si) * NGSW + g0) * SW_REC;
  const bool isfull = s == 1;
  const int toclear = isfull ? -NGSW * SW_REC : 0;                     // FULL -> CLEAR record of the same (tile, level, g)
  uint32_t mw[NG];                                                     // McICA bits of the current 32 layers (FULL stream only)
  const int ku = ws.kslot[s == 0 ? K_CU : s == 1 ? K_FU : s == 2 ? K_NU : K_XU];
  const int kd = ws.kslot[s == 0 ? K_CD : s == 1 ? K_FD : s == 2 ? K_ND : K_XD];
  float *__restrict__ bpart = ws.bpart + (size_t)grp * (nlay + 1) * nk * pcap + c;
  const float *__restrict__ prev = bpart - (size_t)(nlay + 1) * nk * pcap;       // the previous group's running totals (grp > 0)
  const int band = c_sw_grp_band[grp];
  const bool uvband = band >= 9 && band <= 12;

  float zinc[NG], tdbt[NG], tdn[NG], rdnd[NG];
#pragma unroll
  for (int i = 0; i < NG; i++) { zinc[i] = ws.zinc[(size_t)(g0 + i) * pcap + c]; tdbt[i] = 1.f; tdn[i] = 1.f; rdnd[i] = 0.f; }
  for (int lev = nlay; lev >= 0; lev--) {
    const float *__restrict__ q = rec + (size_t)((unsigned)lev * lvstride);      // the group's records: immediate offsets
    float2 R[NG]; float4 P[NG]; float e[NG];
    const int lay = lev - 1;
    if (isfull && lev > 0 && ((lay & 31) == 31 || lev == nlay)) {
#pragma unroll
      for (int i = 0; i < NG; i++) mw[i] = ws.mask[((size_t)(g0 + i) * ws.W + (lay >> 5)) * ws.cap + c];
    }
#pragma unroll
    for (int i = 0; i < NG; i++) {
      R[i] = __ldcs(reinterpret_cast<const float2 *>(q + i * SW_REC + SW_REC_R) + lane);
      if (lev > 0) {
        // the layer's two-stream solution: own record, or the CLEAR stream's where this FULL layer is not cloudy
        const bool own = !isfull || ((mw[i] >> (lay & 31)) & 1u);
        const float *__restrict__ qs = q + i * SW_REC + (own ? 0 : toclear);
        P[i] = __ldcs(reinterpret_cast<const float4 *>(qs) + lane);
        e[i] = __ldcs(qs + SW_REC_E + lane);
      }
    }
    // running sums over ALL g-points in index order, exactly the reference's accumulation (SW:8617-8650): a group starts from
    // the totals the previous group left (the groups' launches follow one another on the stream)
    float su = grp > 0 ? prev[((size_t)lev * nk + ku) * pcap] : 0.f;
    float sd = grp > 0 ? prev[((size_t)lev * nk + kd) * pcap] : 0.f;
    float suv = 0.f, sni = 0.f;
    if (isfull && lev == 0 && grp > 0) { suv = ws.uvni[((size_t)(grp - 1) * 2 + 0) * pcap + c]; sni = ws.uvni[((size_t)(grp - 1) * 2 + 1) * pcap + c]; }
#pragma unroll
    for (int i = 0; i < NG; i++) {
      // flux at interface lev (SW:8036-8045), the reference's operation order and rounding
      const float ru = R[i].x, rud = R[i].y;
      const float zreflect = R_(S_(1.f, M_(rdnd[i], rud)));
      const float dif = S_(tdn[i], tdbt[i]);
      const float fu = M_(A_(M_(tdbt[i], ru), M_(dif, rud)), zreflect);
      const float fd = A_(tdbt[i], M_(A_(dif, M_(M_(tdbt[i], ru), rdnd[i])), zreflect));
      su = A_(su, M_(zinc[i], fu));
      sd = A_(sd, M_(zinc[i], fd));
      if (isfull && lev == 0) { if (uvband) suv = A_(suv, M_(zinc[i], fd)); else sni = A_(sni, M_(zinc[i], fd)); }
    }
    __stcs(bpart + ((size_t)lev * nk + ku) * pcap, su);
    __stcs(bpart + ((size_t)lev * nk + kd) * pcap, sd);
    if (lev == 0) {
      if (isfull) { ws.uvni[((size_t)grp * 2 + 0) * pcap + c] = suv; ws.uvni[((size_t)grp * 2 + 1) * pcap + c] = sni; }
      break;
    }
#pragma unroll
    for (int i = 0; i < NG; i++) {
      // transmittances through the layer below the interface (SW:8007-8034); with tdn = tdbt = 1, rdnd = 0 at the top the
      // general step reproduces the reference's special first step (tdn = tra, rdnd = refd) exactly
      const float zr = R_(S_(1.f, M_(P[i].y, rdnd[i])));
      const float ntdn = A_(M_(tdbt[i], P[i].z), M_(M_(P[i].w, A_(S_(tdn[i], tdbt[i]), M_(M_(tdbt[i], P[i].x), rdnd[i]))), zr));
      const float nrdnd = A_(P[i].y, M_(M_(M_(P[i].w, P[i].w), rdnd[i]), zr));
      tdn[i] = ntdn; rdnd[i] = nrdnd;
      tdbt[i] = M_(e[i], tdbt[i]);
    }
  }
}

int sw_sweep_groups() { return h_sw_grp.n; }
void launch_sw_sweep(const SwArgs &a, cudaStream_t s) {
  const int nstream = 2 + ((a.variants & ARC_VAR_CLEAN) ? 1 : 0) + ((a.variants & ARC_VAR_CLEANCLEAR) ? 1 : 0);
  const int grid = ((a.ncols + 127) / 128) * nstream;
  for (int q = 0; q < h_sw_grp.n; q++) {
    const int g0 = h_sw_grp.g0[q];
    switch (h_sw_grp.ng[q]) {
#define SWEEP_CASE(N) case N: k_sw_sweep<N><<<grid, 128, 0, s>>>(a, q, g0, nstream); break;
      SWEEP_CASE(1) SWEEP_CASE(2) SWEEP_CASE(3) SWEEP_CASE(4) SWEEP_CASE(5) SWEEP_CASE(6) SWEEP_CASE(7) SWEEP_CASE(8)
      SWEEP_CASE(9) SWEEP_CASE(10) SWEEP_CASE(11) SWEEP_CASE(12) SWEEP_CASE(13) SWEEP_CASE(14) SWEEP_CASE(15) SWEEP_CASE(16)
#undef SWEEP_CASE
      default: break;
    }
  }
  count_launch(h_sw_grp.n);
}

// ------------------------------------------------------------------------------------------------------
// Sum of the sweep-group (= band) partials in band order (the g-points inside a band were added in index order by
// k_sw_sweep: the reference's accumulation SW:8617-8650), fluxes -> heating rates (SW:9391-9434) and scatter to the WRF
// arrays (SW:11125-11172).  Block = 64 columns x 4 level-lanes; net fluxes meet in shared memory for the heating rates.
// All partial-buffer reads are coalesced over columns.
constexpr int RED_CX = 64, RED_LY = 4;
__global__ void __launch_bounds__(RED_CX * RED_LY, 4) k_sw_reduce(SwArgs a) {
  __shared__ float s_net[161][RED_CX];
  const int cx = threadIdx.x, ly = threadIdx.y;
  const int c = blockIdx.x * RED_CX + cx;
  const Geo &G = a.geo;
  const SwWs &ws = a.ws;
  const int nlay = ws.nlay, nz = nlay - 1;
  const size_t cap = ws.pcap;      // the reduce only touches the partial buffers
  const bool active = c < a.ncols;
  const bool do_clean = (a.variants & ARC_VAR_CLEAN) != 0;
  const bool do_clnc = (a.variants & ARC_VAR_CLEANCLEAR) != 0;
  int tc = 0, i = 0, j = 0; size_t ij = 0;
  if (active) { tc = ws.cols[c]; G.ij(tc, i, j); ij = G.at2(i, j); }
  for (int lev = ly; lev <= nlay && active; lev += RED_LY) {
    float f[NKIND];
#pragma unroll
    for (int k = 0; k < NKIND; k++) f[k] = 0.f;
    float uvfd = 0.f, nifd = 0.f;
    const int nk = ws.nk;
    // the last sweep group holds the sums over all 112 g-points, accumulated in index order by k_sw_sweep
    const size_t gstride = (size_t)(nlay + 1) * nk * cap;
    const float *p = ws.bpart + (size_t)(a.ngroups - 1) * gstride + ((size_t)lev * nk) * cap + c;
    const unsigned ucap = (unsigned)cap;       // 32-bit kind offsets: nk * pcap < 2^31
    f[K_FU] = p[ws.kslot[K_FU] * ucap]; f[K_FD] = p[ws.kslot[K_FD] * ucap]; f[K_CU] = p[ws.kslot[K_CU] * ucap]; f[K_CD] = p[ws.kslot[K_CD] * ucap];
    if (do_clean) { f[K_NU] = p[ws.kslot[K_NU] * ucap]; f[K_ND] = p[ws.kslot[K_ND] * ucap]; }
    if (do_clnc) { f[K_XU] = p[ws.kslot[K_XU] * ucap]; f[K_XD] = p[ws.kslot[K_XD] * ucap]; }
    if (lev == 0) { uvfd = ws.uvni[((size_t)(a.ngroups - 1) * 2 + 0) * cap + c]; nifd = ws.uvni[((size_t)(a.ngroups - 1) * 2 + 1) * cap + c]; }
    s_net[lev][cx] = f[K_FD] - f[K_FU];
    if (a.swupflx) {        // lev <= nz + 1 always
      const size_t q = G.atp(i, G.kts + lev, j);
      a.swupflx[q] = f[K_FU]; a.swupflxc[q] = f[K_CU]; a.swupflxcln[q] = f[K_NU];
      a.swdnflx[q] = f[K_FD]; a.swdnflxc[q] = f[K_CD]; a.swdnflxcln[q] = f[K_ND];
    }
    if (lev == 0) {
      const float coszrs = a.xcoszen[ij];
      a.gsw[ij] = f[K_FD] - f[K_FU];
      if (a.swupt) {
        a.swupb[ij] = f[K_FU]; a.swupbc[ij] = f[K_CU]; a.swupbcln[ij] = f[K_NU];
        a.swdnb[ij] = f[K_FD]; a.swdnbc[ij] = f[K_CD]; a.swdnbcln[ij] = f[K_ND];
      }
      if (a.swuptclnc) { a.swupbclnc[ij] = f[K_XU]; a.swdnbclnc[ij] = f[K_XD]; }
      // direct / diffuse split at the surface
      float dirall = 0.f, diruv = 0.f, dirni = 0.f;
      for (int g = 0; g < NGSW; g++) {
        const float d = ws.dirs[(size_t)g * cap + c];
        dirall = dirall + d;
        const int b = c_sw_ngb[g];
        if (b >= 9 && b <= 12) diruv = diruv + d; else dirni = dirni + d;
      }
      if (a.swupt) {
        a.swvisdir[ij] = diruv; a.swvisdif[ij] = uvfd - diruv;
        a.swnirdir[ij] = dirni; a.swnirdif[ij] = nifd - dirni;
      }
      a.swddir[ij] = dirall;
      a.swddni[ij] = dirall / coszrs;
      a.swddif[ij] = f[K_FD] - dirall;
    }
    if (lev == nlay) {
      a.swcf[ij] = (f[K_FD] - f[K_FU]) - (f[K_CD] - f[K_CU]);
      if (a.swupt) {
        a.swupt[ij] = f[K_FU]; a.swuptc[ij] = f[K_CU]; a.swuptcln[ij] = f[K_NU];
        a.swdnt[ij] = f[K_FD]; a.swdntc[ij] = f[K_CD]; a.swdntcln[ij] = f[K_ND];
      }
      if (a.swuptclnc) { a.swuptclnc[ij] = f[K_XU]; a.swdntclnc[ij] = f[K_XD]; }
    }
  }
  __syncthreads();
  // heating rate of layer L (1-based) between interfaces L-1 and L; the extra top layer is forced to 0 and not output
  for (int L = 1 + ly; L <= nz && active; L += RED_LY) {
    const int k = G.kts + L - 1;
    const float pdp = a.p8w[G.at3(i, k, j)] / 100.f - a.p8w[G.at3(i, k + 1, j)] / 100.f;
    const float zdpgcp = a.tb.heatfac / pdp;
    const float swhr = (s_net[L][cx] - s_net[L - 1][cx]) * zdpgcp;
    const float tten = swhr / 86400.f;
    a.rthratensw[G.at3(i, k, j)] = tten / a.pi3d[G.at3(i, k, j)];
    if (a.dbg.hr) a.dbg.hr[(size_t)tc * nlay + L - 1] = swhr;
  }
  if (active && ly == 0 && a.dbg.hr) a.dbg.hr[(size_t)tc * nlay + nlay - 1] = 0.f;
}
void launch_sw_reduce(const SwArgs &a, cudaStream_t s) {
  k_sw_reduce<<<(a.ncols + RED_CX - 1) / RED_CX, dim3(RED_CX, RED_LY), 0, s>>>(a);
  count_launch();
}

}  // namespace arc
