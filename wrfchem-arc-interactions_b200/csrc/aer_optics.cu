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

// workspace per (section, point): Chebyshev argument x(ln r) of the wet radius, weight = number * pi r^2 (1/cm), volume
// fractions of the 9 classes
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

__global__ void __launch_bounds__(128) k_aer_prep(AerArgs a, float rmin, float rmax, float xrmin, float xrmax) {
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
    // the series argument depends on the radius only: evaluated here once instead of once per wavelength
    w[(size_t)AWS_R * np] = (2.f * logf(fminf(fmaxf(r, rmin), rmax)) - xrmax - xrmin) / (xrmax - xrmin);
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
constexpr int AER_PTS = 256;                                   // points per block
constexpr int AER_ITEMS = AER_MAXBIN * AER_PTS;                // (section, point) items per block
constexpr int AER_NCELL = (AER_NREFR - 1) * (AER_NREFI - 1);   // refractive-index cells; one more bucket holds the empty items

// Block = 256 (column, level) points x one wavelength.  The 12 coefficient rows an item needs (4 corners of its
// refractive-index cell x 3 quantities) are picked by the item's composition, so neighbouring points read different rows and a
// warp's 16-byte shared-memory loads cost 3 wavefronts each - which is what bounded the first version of this kernel.  Here
// the block's items are first counting-sorted by cell in shared memory (buckets padded to an even length); a thread of the
// evaluation loop then takes TWO neighbouring items of one bucket, so every coefficient it loads feeds both, and a warp
// holds items of one cell (two at a bucket edge): the loads are broadcasts.  A broadcast 16-byte load still costs two
// wavefronts (ncu), hence the pairing.  Each item's arithmetic is unchanged and the section sum runs in section order per
// point afterwards, so the result does not depend on the order inside a bucket.
constexpr int AER_PERM = AER_ITEMS + 2 * AER_NCELL;            // sorted list with the padding slots
constexpr int AER_NONE = 0xFFFF;

__global__ void __launch_bounds__(AER_PTS, 3) k_aer_mie(AerArgs a, AerDev d) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *tab = reinterpret_cast<float *>(smem_raw);
  uint64_t *bar = reinterpret_cast<uint64_t *>(tab + AER_TAB_FLOATS);
  float *it_x = reinterpret_cast<float *>(bar + 2);            // x, then Q_ext
  float *it_t = it_x + AER_ITEMS;                              // t, then Q_sca
  float *it_u = it_t + AER_ITEMS;                              // u, then g
  unsigned short *perm = reinterpret_cast<unsigned short *>(it_u + AER_ITEMS);
  unsigned char *it_cell = reinterpret_cast<unsigned char *>(perm + AER_PERM);
  int *cnt = reinterpret_cast<int *>(it_cell + AER_ITEMS);     // [AER_NCELL + 1] bucket counts
  int *off = cnt + AER_NCELL + 1;                              // [AER_NCELL] bucket starts
  __shared__ int s_total;
  // the 20 wavelengths of a point tile are neighbours in launch order: the tile's workspace lines are read from L2
  const int wl = blockIdx.x % AER_NWL;
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid <= AER_NCELL) cnt[tid] = 0;
  // the wavelength's coefficient table arrives by one bulk copy while phases A and B run; it is waited for before phase C
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar, AER_TAB_FLOATS * 4);
    tma_bulk_g2s(tab, d.coef + (size_t)wl * AER_TAB_FLOATS, AER_TAB_FLOATS * 4, bar);
  }
  const int p = (blockIdx.x / AER_NWL) * AER_PTS + tid;
  const bool live = p < a.npts;
  const size_t np = (size_t)a.npts;
  const int nsec = a.nsec;
  const float r_lo = d.refr_lo[wl], li_lo = d.lnrefi_lo[wl];
  const float r_scale = (float)(AER_NREFR - 1) / (d.refr_hi[wl] - r_lo), li_scale = (float)(AER_NREFI - 1) / (d.lnrefi_hi[wl] - li_lo);
  float nrw[AER_NCLASS], niw[AER_NCLASS];
#pragma unroll
  for (int c = 0; c < AER_NCLASS; c++) { nrw[c] = d.nr[c][wl]; niw[c] = d.ni[c][wl]; }
  float wgt[AER_MAXBIN];
  int cid[AER_MAXBIN], rank[AER_MAXBIN];

  // ---- A: refractive index, cell and interpolation weights of every (section, point) item
#pragma unroll
  for (int s = 0; s < AER_MAXBIN; s++) {
    wgt[s] = 0.f;
    cid[s] = AER_NCELL;
    if (s < nsec && live) {
      const float *w = a.ws + (size_t)s * AER_WS_N * np + p;
      const float weight = w[(size_t)AWS_W * np];
      const float x = w[(size_t)AWS_R * np];
      float refr = 0.f, refi = 0.f;
#pragma unroll
      for (int c = 0; c < AER_NCLASS; c++) {
        const float vf = w[(size_t)(AWS_VF + c) * np];
        refr = fmaf(vf, nrw[c], refr);
        refi = fmaf(vf, niw[c], refi);
      }
      if (weight > 0.f) {
        wgt[s] = weight;
        // bilinear cell in (n_r linear, n_i geometric)
        float tr = (refr - r_lo) * r_scale;
        tr = fminf(fmaxf(tr, 0.f), (float)(AER_NREFR - 1));
        const int ir = min((int)tr, AER_NREFR - 2);
        float ti = (logf(fmaxf(refi, 1.e-30f)) - li_lo) * li_scale;
        ti = fminf(fmaxf(ti, 0.f), (float)(AER_NREFI - 1));
        const int ii = min((int)ti, AER_NREFI - 2);
        cid[s] = ir * (AER_NREFI - 1) + ii;
        const int item = s * AER_PTS + tid;
        it_t[item] = tr - (float)ir;
        it_u[item] = ti - (float)ii;
        it_x[item] = x;
      }
    }
  }
  // rank of every item inside its bucket (the order inside a bucket is arbitrary and does not reach the result)
#pragma unroll
  for (int s = 0; s < AER_MAXBIN; s++) {
    if (s < nsec) {
      it_cell[s * AER_PTS + tid] = (unsigned char)cid[s];
      const unsigned m = __match_any_sync(0xffffffffu, cid[s]);
      const int leader = __ffs(m) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(&cnt[cid[s]], __popc(m));
      rank[s] = __shfl_sync(0xffffffffu, base, leader) + __popc(m & ((1u << lane) - 1u));
    }
  }
  __syncthreads();
  // ---- B: bucket offsets (even; an odd bucket gets one padding slot), then the item list in bucket order
  if (tid < 32) {
    int incl[2], n[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int c = tid + 32 * h;
      n[h] = c < AER_NCELL ? cnt[c] : 0;
      int v = n[h] + (n[h] & 1);
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
      incl[h] = v;
    }
    const int tot0 = __shfl_sync(0xffffffffu, incl[0], 31);
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int c = tid + 32 * h;
      if (c < AER_NCELL) {
        const int start = (h ? tot0 : 0) + incl[h] - (n[h] + (n[h] & 1));
        off[c] = start;
        if (n[h] & 1) perm[start + n[h]] = AER_NONE;
      }
    }
    if (tid == 31) s_total = tot0 + incl[1];
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < AER_MAXBIN; s++)
    if (s < nsec && cid[s] < AER_NCELL) perm[off[cid[s]] + rank[s]] = (unsigned short)(s * AER_PTS + tid);
  __syncthreads();
  mbar_wait(bar, 0);
  // ---- C: the Chebyshev series of the sorted items, two per thread
  const int total = s_total;
  for (int pos = 2 * tid; pos < total; pos += 2 * AER_PTS) {
    const int item0 = perm[pos];
    const int item1raw = perm[pos + 1];
    const int item1 = item1raw == AER_NONE ? item0 : item1raw;        // padding slot: the pair's first item again
    const int cellid = it_cell[item0];
    const int ir = cellid / (AER_NREFI - 1), ii = cellid - ir * (AER_NREFI - 1);
    const float x[2] = {it_x[item0], it_x[item1]};
    // series sum_j c_j T_j(x), c_0 halved (chebev), evaluated at the four refractive-index corners (12 accumulators per
    // item), then blended bilinearly.  The coefficient rows are zero-padded from 50 to 52: the padding needs no special case.
    float a4[2][AER_NQ][4];
#pragma unroll
    for (int e = 0; e < 2; e++)
#pragma unroll
      for (int qn = 0; qn < AER_NQ; qn++) { a4[e][qn][0] = 0.f; a4[e][qn][1] = 0.f; a4[e][qn][2] = 0.f; a4[e][qn][3] = 0.f; }
    const float *cell = tab + (ir * AER_NREFI + ii) * AER_NCOEF_PAD;
    const float x2[2] = {2.f * x[0], 2.f * x[1]};
    float T[2][4];
    auto accumulate = [&](int j4) {
#pragma unroll
      for (int qn = 0; qn < AER_NQ; qn++) {
        const float *base = cell + qn * (AER_NREFR * AER_NREFI * AER_NCOEF_PAD) + j4;
        const float4 c00 = *reinterpret_cast<const float4 *>(base);
        const float4 c01 = *reinterpret_cast<const float4 *>(base + AER_NCOEF_PAD);
        const float4 c10 = *reinterpret_cast<const float4 *>(base + AER_NREFI * AER_NCOEF_PAD);
        const float4 c11 = *reinterpret_cast<const float4 *>(base + (AER_NREFI + 1) * AER_NCOEF_PAD);
#pragma unroll
        for (int e = 0; e < 2; e++) {
          a4[e][qn][0] = fmaf(c00.w, T[e][3], fmaf(c00.z, T[e][2], fmaf(c00.y, T[e][1], fmaf(c00.x, T[e][0], a4[e][qn][0]))));
          a4[e][qn][1] = fmaf(c01.w, T[e][3], fmaf(c01.z, T[e][2], fmaf(c01.y, T[e][1], fmaf(c01.x, T[e][0], a4[e][qn][1]))));
          a4[e][qn][2] = fmaf(c10.w, T[e][3], fmaf(c10.z, T[e][2], fmaf(c10.y, T[e][1], fmaf(c10.x, T[e][0], a4[e][qn][2]))));
          a4[e][qn][3] = fmaf(c11.w, T[e][3], fmaf(c11.z, T[e][2], fmaf(c11.y, T[e][1], fmaf(c11.x, T[e][0], a4[e][qn][3]))));
        }
      }
    };
    // j = 0..3: T_0 halved, T_1 = x
#pragma unroll
    for (int e = 0; e < 2; e++) { T[e][0] = 0.5f; T[e][1] = x[e]; T[e][2] = fmaf(x2[e], x[e], -1.f); T[e][3] = fmaf(x2[e], T[e][2], -x[e]); }
    accumulate(0);
#pragma unroll AER_UNROLL
    for (int j4 = 4; j4 < AER_NCOEF_PAD; j4 += 4) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const float t0 = fmaf(x2[e], T[e][3], -T[e][2]);
        const float t1 = fmaf(x2[e], t0, -T[e][3]);
        const float t2 = fmaf(x2[e], t1, -t0);
        const float t3 = fmaf(x2[e], t2, -t1);
        T[e][0] = t0; T[e][1] = t1; T[e][2] = t2; T[e][3] = t3;
      }
      accumulate(j4);
    }
#pragma unroll
    for (int e = 0; e < 2; e++) {
      if (e == 1 && item1raw == AER_NONE) break;                      // padding slot: nothing to store
      const int item = e == 0 ? item0 : item1;
      const float t = it_t[item], u = it_u[item];
      const float w00 = (1.f - t) * (1.f - u), w10 = t * (1.f - u), w01 = (1.f - t) * u, w11 = t * u;
      float acc[AER_NQ];
#pragma unroll
      for (int qn = 0; qn < AER_NQ; qn++) acc[qn] = w00 * a4[e][qn][0] + w01 * a4[e][qn][1] + w10 * a4[e][qn][2] + w11 * a4[e][qn][3];
      const float pext = expf(acc[0]);
      it_x[item] = pext;
      it_t[item] = fminf(expf(acc[1]), pext);
      it_u[item] = expf(acc[2]);
    }
  }
  __syncthreads();
  // ---- D: section sums of each point, in section order
  if (!live) return;
  float ext = 0.f, sca = 0.f, gsc = 0.f;
#pragma unroll
  for (int s = 0; s < AER_MAXBIN; s++) {
    if (s < nsec && wgt[s] > 0.f) {
      const int item = s * AER_PTS + tid;
      const float weight = wgt[s], pext = it_x[item], pscat = it_t[item], pasm = it_u[item];
      ext += weight * pext;
      sca += weight * pscat;
      gsc += weight * pscat * pasm;
    }
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

static int aer_mie_smem() { return AER_TAB_FLOATS * 4 + 16 + AER_ITEMS * (3 * 4 + 1) + AER_PERM * 2 + (2 * AER_NCELL + 1) * 4 + 16; }

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
  cudaFuncSetAttribute(k_aer_mie, cudaFuncAttributeMaxDynamicSharedMemorySize, aer_mie_smem());
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
  k_aer_prep<<<(a.npts + 127) / 128, 128, 0, s>>>(a, g_D.rmin, g_D.rmax, g_D.xrmin, g_D.xrmax);
  k_aer_mie<<<((a.npts + AER_PTS - 1) / AER_PTS) * AER_NWL, AER_PTS, aer_mie_smem(), s>>>(a, g_D);
  count_launch(2);
  return 0;
}

}  // namespace arc
