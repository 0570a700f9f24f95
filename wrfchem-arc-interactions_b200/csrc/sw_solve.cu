// Shortwave spectral solver: taumol_sw + spcvmc_sw (reftra_sw, vrtqdr_sw) for all call variants in one pass,
// one thread per (column, g-point); plus the g-point reduction / heating-rate / scatter kernel.
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

#ifdef ARC_STRICT
#define RCP(x) (1.0f / (x))
#else
#define RCP(x) __fdividef(1.0f, (x))
#endif

static __constant__ SwBandDesc c_sw[14];
static __constant__ int c_sw_ngb[NGSW];    // band index 0..13 of each SW g-point
void upload_band_descs_sw(const HostTables &T) {
  cudaMemcpyToSymbol(c_sw, T.sw, sizeof(SwBandDesc) * 14);
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
  float specparm = __fdiv_rn(cola, b.speccomb);
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
  if (x <= 0.06f) return 1.f - x + 0.5f * x * x;
  const float tblind = __fdiv_rn(x, __fadd_rn(bpade, x));
  const int itind = (int)__fadd_rn(__fmul_rn(10000.0f, tblind), 0.5f);
  return exp_tbl[itind];
}

// reftra_sw (kmodts = 2, PIFM) for one layer.  Returns (ref, refd, tra, trad).
__device__ __forceinline__ float4 sw_reftra(const float *__restrict__ exp_tbl, float bpade, float zg, float prmuz, float zto1, float zw) {
  const float eps = 1.e-08f, zwcrit = 0.9999995f;
  float4 o;
  const float zg3 = 3.f * zg;
  const float zgamma1 = (8.f - zw * (5.f + zg3)) * 0.25f;
  const float zgamma2 = 3.f * (zw * (1.f - zg)) * 0.25f;
  const float zgamma3 = (2.f - zg3 * prmuz) * 0.25f;
  const float zgamma4 = 1.f - zgamma3;
  const float q = zg * RCP(1.f - zg);
  const float denom = fmaxf((1.f - (1.f - zw) * (q * q)), 1.0E-30f);
  const float zwo = zw * RCP(denom);
  if (zwo >= zwcrit) {
    const float za = zgamma1 * prmuz;
    const float za1 = za - zgamma3;
    const float zgt = zgamma1 * zto1;
    const float ze1 = fminf(__fdiv_rn(zto1, prmuz), 500.f);
    const float ze2 = sw_expt(exp_tbl, ze1, bpade);
    const float r = RCP(1.f + zgt);
    o.x = (zgt - za1 * (1.f - ze2)) * r;
    o.z = 1.f - o.x;
    o.y = zgt * r;
    o.w = 1.f - o.y;
    if (ze2 == 1.0f) { o.x = 0.f; o.z = 1.f; o.y = 0.f; o.w = 1.f; }
  } else {
    const float za1 = zgamma1 * zgamma4 + zgamma2 * zgamma3;
    const float za2 = zgamma1 * zgamma3 + zgamma2 * zgamma4;
    const float zrk = sqrtf(zgamma1 * zgamma1 - zgamma2 * zgamma2);
    const float zrp = zrk * prmuz;
    const float zrp1 = 1.f + zrp, zrm1 = 1.f - zrp, zrk2 = 2.f * zrk;
    const float zrpp = 1.f - zrp * zrp;
    const float zrkg = zrk + zgamma1;
    const float zr1 = zrm1 * (za2 + zrk * zgamma3);
    const float zr2 = zrp1 * (za2 - zrk * zgamma3);
    const float zr3 = zrk2 * (zgamma3 - za2 * prmuz);
    const float zr4 = zrpp * zrkg;
    const float zr5 = zrpp * (zrk - zgamma1);
    const float zt1 = zrp1 * (za1 + zrk * zgamma4);
    const float zt2 = zrm1 * (za1 - zrk * zgamma4);
    const float zt3 = zrk2 * (zgamma4 + za1 * prmuz);
    const float zbeta = (zgamma1 - zrk) * RCP(zrkg);
    const float ze1 = fminf(zrk * zto1, 500.f);
    const float ze2 = fminf(__fdiv_rn(zto1, prmuz), 500.f);
    const float zem1 = sw_expt(exp_tbl, ze1, bpade), zep1 = RCP(zem1);
    const float zem2 = sw_expt(exp_tbl, ze2, bpade), zep2 = RCP(zem2);
    const float zdenr = zr4 * zep1 + zr5 * zem1;
    const float zdent = zr4 * zep1 + zr5 * zem1;   // zt4 = zr4, zt5 = zr5
    if (zdenr >= -eps && zdenr <= eps) { o.x = eps; o.z = zem2; }
    else {
      o.x = zw * (zr1 * zep1 - zr2 * zem1 - zr3 * zem2) * RCP(zdenr);
      o.z = zem2 - zem2 * zw * (zt1 * zep1 - zt2 * zem1 - zt3 * zep2) * RCP(zdent);
    }
    const float zemm = zem1 * zem1;
    const float zdend = RCP((1.f - zbeta * zemm) * zrkg);
    o.y = zgamma2 * (1.f - zemm) * zdend;
    o.w = zrk2 * zem1 * zdend;
  }
  return o;
}

// ------------------------------------------------------------------------------------------------------
template <int NL>
__global__ void __launch_bounds__(256) k_sw_solve(SwArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *s_exp = reinterpret_cast<float *>(smem_raw);                 // 10004 floats
  float *S = s_exp + 10004;                                           // slice
  uint64_t *bar = reinterpret_cast<uint64_t *>(S + SLICE_MAX);

  const int g = blockIdx.y;
  const int b = c_sw_ngb[g];
  const SwBandDesc &D = c_sw[b];
  const int band = b + 16;
  {
    StageReq req[2] = {{s_exp, a.tb.sw_exp, 10004 * 4},
                       {S, a.tb.sw_tab + D.slice_base + (size_t)D.slice_floats * (g - D.g0), (uint32_t)D.slice_floats * 4}};
    stage_tables(bar, req, 2);
  }
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.ncols) return;

  const SwWs &ws = a.ws;
  const int nlay = ws.nlay;
  const size_t cap = ws.cap;
  const float bpade = a.tb.bpade, oneminus = a.tb.oneminus;
  const bool do_clean = (a.variants & ARC_VAR_CLEAN) != 0;
  const bool do_clnc = (a.variants & ARC_VAR_CLEANCLEAR) != 0;
  const bool noaer = do_clean || do_clnc;

  const float prmu0 = ws.colf[(size_t)SWF_MU0 * cap + c];
  const bool uv = (b >= 9 && b <= 12);
  const float albp = ws.colf[(size_t)(uv ? SWF_ALBDIR_UV : SWF_ALBDIR_NIR) * cap + c];
  const float albd = ws.colf[(size_t)(uv ? SWF_ALBDIF_UV : SWF_ALBDIF_NIR) * cap + c];
  const int laytrop = ws.laytrop[c];
  const int laysol = ws.laysol[(size_t)b * cap + c];

  uint32_t mw[NL / 32];
#pragma unroll
  for (int w = 0; w < NL / 32; w++) mw[w] = w < ws.W ? ws.mask[((size_t)g * ws.W + w) * cap + c] : 0u;

  // per-layer two-stream solutions and the upward reflectances at every interface (thread-private, coalesced local memory)
  float4 Pa[NL], Pn[NL], Pca[NL], Pcn[NL];
  float Ea[NL], En[NL], Eca[NL], Ecn[NL];
  float2 Ru[4][NL];          // (rup, rupd) at interface lay+1 for streams clear, full, clean, cleanclear

  float rup[4], rupd[4];
#pragma unroll
  for (int s = 0; s < 4; s++) { rup[s] = albp; rupd[s] = albd; }
  float sfluxzen = 0.f;
  float tdir_nodel = 1.f;     // product of the un-delta-scaled direct transmittances of the FULL stream

  const float *coef = ws.coef + c;
  for (int lay = 0; lay < nlay; lay++) {
    SwLay L;
    {
      const size_t st = (size_t)nlay * cap;
      const float *p = coef + (size_t)lay * cap;
      L.fac00 = p[SWC_FAC00 * st]; L.fac01 = p[SWC_FAC01 * st]; L.fac10 = p[SWC_FAC10 * st]; L.fac11 = p[SWC_FAC11 * st];
      L.h2o = p[SWC_H2O * st]; L.co2 = p[SWC_CO2 * st]; L.o3 = p[SWC_O3 * st]; L.ch4 = p[SWC_CH4 * st]; L.o2 = p[SWC_O2 * st];
      L.mol = p[SWC_MOL * st];
      L.selffac = p[SWC_SELFFAC * st]; L.selffrac = p[SWC_SELFFRAC * st]; L.forfac = p[SWC_FORFAC * st]; L.forfrac = p[SWC_FORFRAC * st];
      const int pk = __float_as_int(p[SWC_IDX * st]);
      L.jp = IDX_JP(pk); L.jt = IDX_JT(pk); L.jt1 = IDX_JT1(pk); L.indself = IDX_SELF(pk); L.indfor = IDX_FOR(pk);
    }
    const bool lower = lay < laytrop;
    float taug, taur, sfl = 0.f;
    sw_taumol(band, S, D, L, lower, oneminus, taug, taur, sfl);
    if (lay == laysol) sfluxzen = sfl;
    if (a.dbg.taug) {
      const size_t q = ((size_t)ws.cols[c] * nlay + lay) * NGSW + g;
      a.dbg.taug[q] = taug; a.dbg.taur[q] = taur;
    }
    const float taua = ws.aer[(((size_t)b * 3 + 0) * nlay + lay) * cap + c];
    const float omga = ws.aer[(((size_t)b * 3 + 1) * nlay + lay) * cap + c];
    const float asya = ws.aer[(((size_t)b * 3 + 2) * nlay + lay) * cap + c];
    const bool cloudy = (mw[lay >> 5] >> (lay & 31)) & 1u;
    float taucmc = 0.f, ssacmc = 1.f, asmcmc = 0.f, taormc = 0.f;
    if (cloudy) {
      taucmc = ws.cld[(((size_t)b * 4 + 0) * nlay + lay) * cap + c];
      ssacmc = ws.cld[(((size_t)b * 4 + 1) * nlay + lay) * cap + c];
      asmcmc = ws.cld[(((size_t)b * 4 + 2) * nlay + lay) * cap + c];
      taormc = ws.cld[(((size_t)b * 4 + 3) * nlay + lay) * cap + c];
    }
    if (a.dbg.taucmc) a.dbg.taucmc[((size_t)ws.cols[c] * nlay + lay) * NGSW + g] = taucmc;

    // ---- optical properties of the clear and cloudy layer, with (v = 0) and without (v = 1) aerosol
    float4 pclr[2], pcld[2];
    float eclr[2], ecld[2];
#pragma unroll
    for (int v = 0; v < 2; v++) {
      if (v == 1 && !noaer) break;
      const float ta = v == 0 ? taua : 0.f;
      float ztauc = taur + taug + ta;
      float zomcc = taur * 1.0f + ta * omga;
      float zgcc = asya * omga * ta * RCP(zomcc);
      zomcc = zomcc * RCP(ztauc);
      if (v == 0) {
        // direct beam without delta scaling (diagnostic surface direct flux of the FULL stream)
        const float tauorig = cloudy ? ztauc + taormc : ztauc;
        tdir_nodel = tdir_nodel * sw_expt(s_exp, __fdiv_rn(tauorig, prmu0), bpade);
      }
      const float zf = zgcc * zgcc;
      const float zwf = zomcc * zf;
      ztauc = (1.0f - zwf) * ztauc;
      zomcc = (zomcc - zwf) * RCP(fmaxf(1.0f - zwf, 1.0E-30f));
      zgcc = (zgcc - zf) * RCP(fmaxf(1.0f - zf, 1.0E-30f));
      pclr[v] = sw_reftra(s_exp, bpade, zgcc, prmu0, ztauc, zomcc);
      eclr[v] = sw_expt(s_exp, __fdiv_rn(ztauc, prmu0), bpade);
      if (cloudy) {
        const float ztauo = ztauc + taucmc;
        float zomco = ztauc * zomcc + taucmc * ssacmc;
        const float zgco = (taucmc * ssacmc * asmcmc + ztauc * zomcc * zgcc) * RCP(zomco);
        zomco = zomco * RCP(ztauo);
        pcld[v] = sw_reftra(s_exp, bpade, zgco, prmu0, ztauo, zomco);
        ecld[v] = sw_expt(s_exp, __fdiv_rn(ztauo, prmu0), bpade);
      }
    }
    Pa[lay] = pclr[0]; Ea[lay] = eclr[0];
    if (noaer) { Pn[lay] = pclr[1]; En[lay] = eclr[1]; }
    if (cloudy) {
      Pca[lay] = pcld[0]; Eca[lay] = ecld[0];
      if (do_clean) { Pcn[lay] = pcld[1]; Ecn[lay] = ecld[1]; }
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
      const float zreflect = RCP(1.f - rupd[s] * P.y);
      const float nrup = P.x + (P.w * ((P.z - e) * rupd[s] + e * rup[s])) * zreflect;
      const float nrupd = P.y + P.w * P.w * rupd[s] * zreflect;
      rup[s] = nrup; rupd[s] = nrupd;
      Ru[s][lay] = make_float2(nrup, nrupd);
    }
  }
  if (a.dbg.sfluxzen) a.dbg.sfluxzen[(size_t)ws.cols[c] * NGSW + g] = sfluxzen;

  // ---- top-down sweep: transmittances and fluxes at every interface
  const float zincflx = ws.colf[(size_t)SWF_ADJFLUX * cap + c] * sfluxzen * prmu0;
  float tdbt[4], tdn[4], rdnd[4];
#pragma unroll
  for (int s = 0; s < 4; s++) { tdbt[s] = 1.f; tdn[s] = 1.f; rdnd[s] = 0.f; }
  float *part = ws.part + ((size_t)g * (nlay + 1)) * NKIND * cap + c;
  for (int lev = nlay; lev >= 0; lev--) {
    // flux at interface lev
#pragma unroll
    for (int s = 0; s < 4; s++) {
      if (s == 2 && !do_clean) continue;
      if (s == 3 && !do_clnc) continue;
      float ru, rud;
      if (lev > 0) { const float2 r = Ru[s][lev - 1]; ru = r.x; rud = r.y; } else { ru = albp; rud = albd; }
      const float zreflect = RCP(1.f - rdnd[s] * rud);
      const float fu = (tdbt[s] * ru + (tdn[s] - tdbt[s]) * rud) * zreflect;
      const float fd = tdbt[s] + (tdn[s] - tdbt[s] + tdbt[s] * ru * rdnd[s]) * zreflect;
      const int ku = s == 0 ? K_CU : s == 1 ? K_FU : s == 2 ? K_NU : K_XU;
      part[((size_t)lev * NKIND + ku) * cap] = zincflx * fu;
      part[((size_t)lev * NKIND + ku + 1) * cap] = zincflx * fd;
    }
    if (lev == 0) break;
    const int lay = lev - 1;
    const bool cloudy = (mw[lay >> 5] >> (lay & 31)) & 1u;
    float4 pa = Pa[lay]; float ea = Ea[lay];
    float4 pn = pa; float en = ea;
    if (noaer) { pn = Pn[lay]; en = En[lay]; }
    float4 pf = pa, pc = pn; float ef = ea, ec = en;
    if (cloudy) { pf = Pca[lay]; ef = Eca[lay]; if (do_clean) { pc = Pcn[lay]; ec = Ecn[lay]; } }
#pragma unroll
    for (int s = 0; s < 4; s++) {
      if (s == 2 && !do_clean) continue;
      if (s == 3 && !do_clnc) continue;
      const float4 P = s == 0 ? pa : s == 1 ? pf : s == 2 ? pc : pn;
      const float e = s == 0 ? ea : s == 1 ? ef : s == 2 ? ec : en;
      const float zreflect = RCP(1.f - P.y * rdnd[s]);
      const float ntdn = tdbt[s] * P.z + (P.w * ((tdn[s] - tdbt[s]) + tdbt[s] * P.x * rdnd[s])) * zreflect;
      const float nrdnd = P.y + P.w * P.w * rdnd[s] * zreflect;
      tdn[s] = ntdn; rdnd[s] = nrdnd;
      tdbt[s] = e * tdbt[s];
    }
  }
  ws.dirs[(size_t)g * cap + c] = zincflx * tdir_nodel;
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
  dim3 grid((a.ncols + 255) / 256, NGSW);
  if (a.ws.nlay <= 64) k_sw_solve<64><<<grid, 256, sw_solve_smem(), s>>>(a);
  else if (a.ws.nlay <= 128) k_sw_solve<128><<<grid, 256, sw_solve_smem(), s>>>(a);
  else k_sw_solve<160><<<grid, 256, sw_solve_smem(), s>>>(a);
  count_launch();
}

// ------------------------------------------------------------------------------------------------------
// Reduction over g-points (in index order = the reference's accumulation order SW:8617-8650), fluxes -> heating
// rates (SW:9391-9434) and scatter to the WRF arrays (SW:11125-11172).  One thread per column; all partial-buffer
// reads are coalesced over columns.
__global__ void __launch_bounds__(128) k_sw_reduce(SwArgs a) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.ncols) return;
  const Geo &G = a.geo;
  const SwWs &ws = a.ws;
  const int nlay = ws.nlay, nz = nlay - 1;
  const size_t cap = ws.cap;
  const int tc = ws.cols[c];
  int i, j; G.ij(tc, i, j);
  const size_t ij = G.at2(i, j);
  const bool do_clean = (a.variants & ARC_VAR_CLEAN) != 0;
  const bool do_clnc = (a.variants & ARC_VAR_CLEANCLEAR) != 0;
  const float coszrs = a.xcoszen[ij];

  float net_prev = 0.f;
  for (int lev = 0; lev <= nlay; lev++) {
    float f[NKIND];
#pragma unroll
    for (int k = 0; k < NKIND; k++) f[k] = 0.f;
    float uvfd = 0.f, nifd = 0.f;
    for (int g = 0; g < NGSW; g++) {
      const float *p = ws.part + (((size_t)g * (nlay + 1) + lev) * NKIND) * cap + c;
      f[K_FU] = f[K_FU] + p[(size_t)K_FU * cap];
      const float fd = p[(size_t)K_FD * cap];
      f[K_FD] = f[K_FD] + fd;
      f[K_CU] = f[K_CU] + p[(size_t)K_CU * cap];
      f[K_CD] = f[K_CD] + p[(size_t)K_CD * cap];
      if (do_clean) { f[K_NU] = f[K_NU] + p[(size_t)K_NU * cap]; f[K_ND] = f[K_ND] + p[(size_t)K_ND * cap]; }
      if (do_clnc) { f[K_XU] = f[K_XU] + p[(size_t)K_XU * cap]; f[K_XD] = f[K_XD] + p[(size_t)K_XD * cap]; }
      if (lev == 0) {
        const int b = c_sw_ngb[g];
        if (b >= 9 && b <= 12) uvfd = uvfd + fd; else nifd = nifd + fd;
      }
    }
    // heating rate of the layer below this interface
    const float net = f[K_FD] - f[K_FU];
    if (lev >= 1 && lev <= nz) {
      const int k = G.kts + lev - 1;
      const float pdp = a.p8w[G.at3(i, k, j)] / 100.f - a.p8w[G.at3(i, k + 1, j)] / 100.f;
      const float zdpgcp = a.tb.heatfac / pdp;
      const float swhr = (net - net_prev) * zdpgcp;
      const float tten = swhr / 86400.f;
      a.rthratensw[G.at3(i, k, j)] = tten / a.pi3d[G.at3(i, k, j)];
      if (a.dbg.hr) a.dbg.hr[(size_t)tc * nlay + lev - 1] = swhr;
    }
    net_prev = net;
    if (lev <= nz + 1 && a.swupflx) {
      const size_t q = G.atp(i, G.kts + lev, j);
      a.swupflx[q] = f[K_FU]; a.swupflxc[q] = f[K_CU]; a.swupflxcln[q] = f[K_NU];
      a.swdnflx[q] = f[K_FD]; a.swdnflxc[q] = f[K_CD]; a.swdnflxcln[q] = f[K_ND];
    }
    if (lev == 0) {
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
  if (a.dbg.hr) a.dbg.hr[(size_t)tc * nlay + nlay - 1] = 0.f;
}
void launch_sw_reduce(const SwArgs &a, cudaStream_t s) {
  k_sw_reduce<<<(a.ncols + 127) / 128, 128, 0, s>>>(a);
  count_launch();
}

}  // namespace arc
