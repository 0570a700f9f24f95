// Aerosol optical properties: optical_prep_sectional / optical_prep_modal + mieaer (Chebyshev-Mie) on the GPU.
//
// Upstream algorithm: WRF-Chem v3.9.1 chem/module_optical_averaging.F.  That file is NOT in the reference repository
// (SURVEY.md section 0.4); only its outputs are consumed there (module_radiation_driver.F:113-124, registry.chem:1332-1390).
// This is a restatement of the published algorithm and is "self-consistent only": parity is pinned against our own CPU
// restatement (the test oracle) and against direct Mie theory, not against the Fortran.
//
//   k_aer_prep   one thread per (column, level): per size section, species volumes (mass / density), volume-averaged
//                complex refractive index weights, wet radius and number; modal input is first mapped onto the
//                sections with the log-normal error-function integrals (optical_prep_modal).
//   k_aer_mie    one block per wavelength x 256 (column, level) points: the wavelength's Chebyshev coefficient table
//                (7 x 7 refractive indices x 3 quantities x 50 coefficients, 30 KB) is staged in shared memory by one TMA
//                bulk copy; each thread loops over the sections: bilinear weights in (n_r, ln n_i), Chebyshev
//                polynomials in ln r, Q_ext, Q_sca, g; section sums -> tau, omega, g (SW) and absorption tau (LW).
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/arc_rad.h"
#include "aer.h"

namespace arc {

// workspace per (section, point): radius (cm), weight = number * pi r^2 (1/cm), volume fractions of the 9 classes
enum { AWS_R = 0, AWS_W, AWS_VF, AER_WS_N = 2 + AER_NCLASS };

__constant__ float c_dens[AER_NCLASS] = {1.8f, 1.8f, 2.2f, 1.8f, 2.2f, 2.6f, 1.0f, 1.7f, 1.0f};   // g/cm3: so4 no3 cl nh4 na oin oc bc water

__device__ inline void point_ijk(const Geo &G, int p, int &i, int &k, int &j) {
  const int nz = G.kte - G.kts + 1;
  const int tc = p / nz;
  k = G.kts + p % nz;
  G.ij(tc, i, j);
}

// section edges (cm): MOSAIC sections, 3.90625e-6 .. 1.0e-3 cm log-spaced (module_mosaic_driver.F:6052-6069 of the v3.6.1 tree)
__device__ inline void section_edges(int nsec, int s, float &dlo, float &dhi) {
  const float lo = 3.90625e-6f, hi = 1.0e-3f;
  const float r = logf(hi / lo) / (float)nsec;
  dlo = lo * expf(r * (float)s);
  dhi = lo * expf(r * (float)(s + 1));
}

__global__ void __launch_bounds__(128) k_aer_prep(AerArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.npts) return;
  const AerSpecList &sl = *a.sl;
  int i, k, j; point_ijk(a.geo, p, i, k, j);
  const size_t q = a.geo.at3(i, k, j);
  const float rho = 1.0f / a.alt[q];                 // kg/m3
  const float conv_m = rho * 1.0e-12f;               // ug/kg -> g/cm3(air)
  const float conv_n = rho * 1.0e-6f;                // #/kg  -> #/cm3(air)
  const int nsec = a.nsec;
  const size_t np = (size_t)a.npts;
  float vol[AER_MAXBIN][AER_NCLASS], num[AER_MAXBIN];
  for (int s = 0; s < nsec; s++) { num[s] = 0.f; for (int c = 0; c < AER_NCLASS; c++) vol[s][c] = 0.f; }
  if (sl.mode == 1) {
    // optical_prep_sectional: volumes by class, number
    for (int s = 0; s < nsec; s++) {
      for (int m = 0; m < sl.nspec[s]; m++) {
        const int c = sl.cls[s][m];
        vol[s][c] += fmaxf(sl.mass[s][m][q], 0.f) * conv_m / c_dens[c];
      }
      num[s] = fmaxf(sl.num[s][q], 0.f) * conv_n;
    }
  } else {
    // optical_prep_modal: each log-normal mode is distributed over the sections
    for (int md = 0; md < sl.nbin; md++) {
      float vm[AER_NCLASS];
      for (int c = 0; c < AER_NCLASS; c++) vm[c] = 0.f;
      float vtot = 0.f;
      for (int m = 0; m < sl.nspec[md]; m++) {
        const int c = sl.cls[md][m];
        const float v = fmaxf(sl.mass[md][m][q], 0.f) * conv_m / c_dens[c];
        vm[c] += v; vtot += v;
      }
      const float nm = fmaxf(sl.num[md][q], 0.f) * conv_n;
      if (!(vtot > 1.e-30f) || !(nm > 1.e-20f)) continue;
      // the section fractions are differences of error functions close to 0 or 1 in the tails: double precision
      const double lns = log((double)sl.sigmag[md]);
      // V = pi/6 N dg^3 exp(4.5 ln^2 sigma)
      const double dgn = cbrt((double)vtot / (0.5235987755982988 * (double)nm)) * exp(-1.5 * lns * lns);
      const double dgv = dgn * exp(3.0 * lns * lns);
      const double rs2 = 0.7071067811865476 / lns;
      double fn_prev = 0.0, fv_prev = 0.0;
      const double lo = 3.90625e-6, lr = log(1.0e-3 / lo) / (double)nsec;
      for (int s = 0; s < nsec; s++) {
        const double dhi = lo * exp(lr * (double)(s + 1));
        // cumulative fractions below dhi; the highest section also takes the upper tail
        const double fn = s == nsec - 1 ? 1.0 : 0.5 * (1.0 + erf(log(dhi / dgn) * rs2));
        const double fv = s == nsec - 1 ? 1.0 : 0.5 * (1.0 + erf(log(dhi / dgv) * rs2));
        num[s] += (float)((double)nm * (fn - fn_prev));
        for (int c = 0; c < AER_NCLASS; c++) vol[s][c] += (float)((double)vm[c] * (fv - fv_prev));
        fn_prev = fn; fv_prev = fv;
      }
    }
  }
  for (int s = 0; s < nsec; s++) {
    float vdry = 0.f;
    for (int c = 0; c < AER_NCLASS - 1; c++) vdry += vol[s][c];
    const float vwet = vdry + vol[s][AER_NCLASS - 1];
    float *w = a.ws + (size_t)s * AER_WS_N * np + p;
    if (!(vdry > 1.e-30f)) {
      w[(size_t)AWS_R * np] = 0.f; w[(size_t)AWS_W * np] = 0.f;
      for (int c = 0; c < AER_NCLASS; c++) w[(size_t)(AWS_VF + c) * np] = 0.f;
      continue;
    }
    float dlo, dhi; section_edges(nsec, s, dlo, dhi);
    float n = num[s];
    float dp_dry = n > 1.e-20f ? cbrtf(1.9098593f * vdry / n) : 0.f;
    if (!(dp_dry >= dlo && dp_dry <= dhi)) {          // outside the section: centre diameter, number from the volume
      dp_dry = sqrtf(dlo * dhi);
      n = 1.9098593f * vdry / (dp_dry * dp_dry * dp_dry);
    }
    const float dp_wet = dp_dry * cbrtf(vwet / vdry);
    const float r = 0.5f * dp_wet;
    w[(size_t)AWS_R * np] = r;
    w[(size_t)AWS_W * np] = n * 3.14159265f * r * r;
    const float inv = 1.0f / vwet;
    for (int c = 0; c < AER_NCLASS; c++) w[(size_t)(AWS_VF + c) * np] = vol[s][c] * inv;
  }
}

#ifndef AER_UNROLL_J
#define AER_UNROLL_J 1
#endif
constexpr int AER_UNROLL = AER_UNROLL_J;
constexpr int AER_TAB_FLOATS = AER_NQ * AER_NREFR * AER_NREFI * AER_NCOEF_PAD;     // 7644 floats = 30,576 B per wavelength

__global__ void __launch_bounds__(256) k_aer_mie(AerArgs a, AerDev d) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *tab = reinterpret_cast<float *>(smem_raw);
  uint64_t *bar = reinterpret_cast<uint64_t *>(tab + AER_TAB_FLOATS);
  const int wl = blockIdx.y;
  {
    StageReq req[1] = {{tab, d.coef + (size_t)wl * AER_TAB_FLOATS, AER_TAB_FLOATS * 4}};
    stage_tables(bar, req, 1);
  }
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.npts) return;
  const size_t np = (size_t)a.npts;
  float ext = 0.f, sca = 0.f, gsc = 0.f;
  const float xrmin = d.xrmin, xrmax = d.xrmax;
  const float r_lo = d.refr_lo[wl], r_hi = d.refr_hi[wl], li_lo = d.lnrefi_lo[wl], li_hi = d.lnrefi_hi[wl];
  for (int s = 0; s < a.nsec; s++) {
    const float *w = a.ws + (size_t)s * AER_WS_N * np + p;
    const float weight = w[(size_t)AWS_W * np];
    if (!(weight > 0.f)) continue;
    float r = w[(size_t)AWS_R * np];
    r = fminf(fmaxf(r, d.rmin), d.rmax);
    float refr = 0.f, refi = 0.f;
#pragma unroll
    for (int c = 0; c < AER_NCLASS; c++) {
      const float vf = w[(size_t)(AWS_VF + c) * np];
      refr = fmaf(vf, d.nr[c][wl], refr);
      refi = fmaf(vf, d.ni[c][wl], refi);
    }
    // bilinear cell in (n_r linear, n_i geometric)
    float tr = (refr - r_lo) / (r_hi - r_lo) * (float)(AER_NREFR - 1);
    tr = fminf(fmaxf(tr, 0.f), (float)(AER_NREFR - 1));
    int ir = min((int)tr, AER_NREFR - 2);
    const float t = tr - (float)ir;
    float ti = (logf(fmaxf(refi, 1.e-30f)) - li_lo) / (li_hi - li_lo) * (float)(AER_NREFI - 1);
    ti = fminf(fmaxf(ti, 0.f), (float)(AER_NREFI - 1));
    int ii = min((int)ti, AER_NREFI - 2);
    const float u = ti - (float)ii;
    const float w00 = (1.f - t) * (1.f - u), w10 = t * (1.f - u), w01 = (1.f - t) * u, w11 = t * u;
    const float x = (2.f * logf(r) - xrmax - xrmin) / (xrmax - xrmin);
    // series sum_j c_j T_j(x), c_0 halved (chebev), evaluated at the four refractive-index corners (12 accumulators), then
    // blended bilinearly.  The coefficient rows are zero-padded from 50 to 52, so the padding needs no special case.
    float a4[AER_NQ][4];
#pragma unroll
    for (int qn = 0; qn < AER_NQ; qn++) { a4[qn][0] = 0.f; a4[qn][1] = 0.f; a4[qn][2] = 0.f; a4[qn][3] = 0.f; }
    const float *cell = tab + (ir * AER_NREFI + ii) * AER_NCOEF_PAD;
    float tjm1 = 1.f, tj = x;          // T_0, T_1
    const float x2 = 2.f * x;
#pragma unroll AER_UNROLL
    for (int j4 = 0; j4 < AER_NCOEF_PAD; j4 += 4) {
      float T0, T1, T2, T3;
      if (j4 == 0) {
        T0 = 0.5f; T1 = x;
        T2 = fmaf(x2, tj, -tjm1); T3 = fmaf(x2, T2, -tj);
      } else {
        T0 = fmaf(x2, tj, -tjm1); T1 = fmaf(x2, T0, -tj); T2 = fmaf(x2, T1, -T0); T3 = fmaf(x2, T2, -T1);
      }
      tjm1 = T2; tj = T3;
#pragma unroll
      for (int qn = 0; qn < AER_NQ; qn++) {
        const float *base = cell + qn * (AER_NREFR * AER_NREFI * AER_NCOEF_PAD) + j4;
        const float4 c00 = *reinterpret_cast<const float4 *>(base);
        const float4 c01 = *reinterpret_cast<const float4 *>(base + AER_NCOEF_PAD);
        const float4 c10 = *reinterpret_cast<const float4 *>(base + AER_NREFI * AER_NCOEF_PAD);
        const float4 c11 = *reinterpret_cast<const float4 *>(base + (AER_NREFI + 1) * AER_NCOEF_PAD);
        a4[qn][0] = fmaf(c00.w, T3, fmaf(c00.z, T2, fmaf(c00.y, T1, fmaf(c00.x, T0, a4[qn][0]))));
        a4[qn][1] = fmaf(c01.w, T3, fmaf(c01.z, T2, fmaf(c01.y, T1, fmaf(c01.x, T0, a4[qn][1]))));
        a4[qn][2] = fmaf(c10.w, T3, fmaf(c10.z, T2, fmaf(c10.y, T1, fmaf(c10.x, T0, a4[qn][2]))));
        a4[qn][3] = fmaf(c11.w, T3, fmaf(c11.z, T2, fmaf(c11.y, T1, fmaf(c11.x, T0, a4[qn][3]))));
      }
    }
    float acc[AER_NQ];
#pragma unroll
    for (int qn = 0; qn < AER_NQ; qn++) acc[qn] = w00 * a4[qn][0] + w01 * a4[qn][1] + w10 * a4[qn][2] + w11 * a4[qn][3];
    const float pext = expf(acc[0]);
    const float pscat = fminf(expf(acc[1]), pext);
    const float pasm = expf(acc[2]);
    ext += weight * pext;
    sca += weight * pscat;
    gsc += weight * pscat * pasm;
  }
  int i, k, j; point_ijk(a.geo, p, i, k, j);
  const size_t q = a.geo.at3(i, k, j);
  const float dzcm = a.dz8w[q] * 100.f;
  if (wl < AER_NSW) {
    a.tauaer[wl][q] = ext * dzcm;
    a.waer[wl][q] = ext > 0.f ? sca / ext : 1.f;
    a.gaer[wl][q] = sca > 0.f ? gsc / sca : 0.f;
  } else {
    const float absb = fmaxf(ext - sca, 0.f);              // RRTMG_LW is absorption-only
    a.tauaerlw[wl - AER_NSW][q] = absb * dzcm;
    if (a.extaerlw[wl - AER_NSW]) a.extaerlw[wl - AER_NSW][q] = absb * 1.0e5f;       // 1/cm -> 1/km
  }
}

// ---- host side --------------------------------------------------------------------------------------------------
static AerTables g_T;
static AerDev g_D;
static bool g_aer_ready = false;
static float *g_coef = nullptr;
static AerSpecList *g_sl = nullptr;
static float *g_ws = nullptr; static size_t g_ws_bytes = 0;

int aer_init(const float *nr, const float *ni, std::string &err) {
  float dnr[AER_NCLASS][AER_NWL], dni[AER_NCLASS][AER_NWL];
  default_refindex(dnr, dni);
  if (nr && ni) { memcpy(dnr, nr, sizeof(dnr)); memcpy(dni, ni, sizeof(dni)); }
  build_aer_tables(dnr, dni, g_T);
  if (g_coef) cudaFree(g_coef);
  if (cudaMalloc(&g_coef, g_T.coef.size() * 4) != cudaSuccess) { err = "aer_init: cudaMalloc failed"; return ARC_ERR_CUDA; }
  cudaMemcpy(g_coef, g_T.coef.data(), g_T.coef.size() * 4, cudaMemcpyHostToDevice);
  if (!g_sl && cudaMalloc(&g_sl, sizeof(AerSpecList)) != cudaSuccess) { err = "aer_init: cudaMalloc failed"; return ARC_ERR_CUDA; }
  g_D.coef = g_coef;
  memcpy(g_D.nr, g_T.nr, sizeof(g_D.nr)); memcpy(g_D.ni, g_T.ni, sizeof(g_D.ni));
  for (int w = 0; w < AER_NWL; w++) {
    g_D.refr_lo[w] = g_T.refr_lo[w]; g_D.refr_hi[w] = g_T.refr_hi[w];
    g_D.lnrefi_lo[w] = logf(g_T.refi_lo[w]); g_D.lnrefi_hi[w] = logf(g_T.refi_hi[w]);
  }
  g_D.rmin = (float)g_T.rmin; g_D.rmax = (float)g_T.rmax;
  g_D.xrmin = logf(g_D.rmin); g_D.xrmax = logf(g_D.rmax);
  cudaFuncSetAttribute(k_aer_mie, cudaFuncAttributeMaxDynamicSharedMemorySize, AER_TAB_FLOATS * 4 + 16);
  g_aer_ready = true;
  return 0;
}
bool aer_ready() { return g_aer_ready; }
void aer_finalize() {
  if (g_coef) cudaFree(g_coef);
  if (g_sl) cudaFree(g_sl);
  if (g_ws) cudaFree(g_ws);
  g_coef = nullptr; g_sl = nullptr; g_ws = nullptr; g_ws_bytes = 0; g_aer_ready = false;
}
const AerTables &aer_tables() { return g_T; }

int aer_run(AerArgs &a, const AerSpecList &sl, cudaStream_t s, std::string &err) {
  const size_t need = (size_t)a.nsec * AER_WS_N * a.npts * 4;
  if (need > g_ws_bytes) {
    if (g_ws) cudaFree(g_ws);
    g_ws = nullptr; g_ws_bytes = 0;
    if (cudaMalloc(&g_ws, need) != cudaSuccess) { err = "arc_aer_optics: workspace allocation failed"; return ARC_ERR_CUDA; }
    g_ws_bytes = need;
  }
  a.ws = g_ws;
  a.sl = g_sl;
  cudaMemcpyAsync(g_sl, &sl, sizeof(AerSpecList), cudaMemcpyHostToDevice, s);
  k_aer_prep<<<(a.npts + 127) / 128, 128, 0, s>>>(a);
  dim3 grid((a.npts + 255) / 256, AER_NWL);
  k_aer_mie<<<grid, 256, AER_TAB_FLOATS * 4 + 16, s>>>(a, g_D);
  count_launch(2);
  return 0;
}

}  // namespace arc
