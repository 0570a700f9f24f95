// Column preparation kernels: sunlit-column compaction, McICA sub-column masks, and the
// WRF->RRTMG adapter + inatm + setcoef + aerosol/cloud band optics, one thread per column, writing
// the [field][layer][column] workspace the spectral solvers read.
//
// Compiled with -fmad=false: jp/jt/jt1/indfor/indself/indminor/laytrop and the McICA masks must be
// bit-identical to the reference's unfused arithmetic.
//
// Reference (module_ra_rrtmg_sw.F = SW, module_ra_rrtmg_lw.F = LW, v3.9.1):
//   RRTMG_SWRAD SW:10320-11071, inatm_sw SW:9520-9873, setcoef_sw SW:2734-2990, cldprmc_sw SW:1969-2390,
//   mcica_subcol_sw/generate_stochastic_clouds_sw/kissvec SW:1392-1932 (LW twins LW:2089-2618),
//   RRTMG_LWRAD LW:11877-12629, inatm LW:11067-11403, setcoef LW:3444-3809, cldprmc LW:2653-2914,
//   taumol_sw's laysolfr selection SW:3293-4538.
#include <vector>

#include "adapter.cuh"
#include "../../include/arc_rad.h"

namespace arc {

static __constant__ SwBandDesc c_sw[14];

static long long g_launches = 0;
void count_launch(int n) { g_launches += n; }
long long launch_count() { return g_launches; }

void upload_band_descs_sw(const HostTables &T);   // sw_solve.cu
void upload_band_descs_lw(const HostTables &T);   // lw_solve.cu
void upload_band_descs(const HostTables &T) {
  cudaMemcpyToSymbol(c_sw, T.sw, sizeof(SwBandDesc) * 14);
  upload_band_descs_sw(T);
  upload_band_descs_lw(T);
}

__device__ __forceinline__ void set_status(int *status, int code) {
  if (code) atomicCAS(status, 0, code);
}

// ------------------------------------------------------------------------------------------------------
// Sunlit compaction (SW:10336 "if (coszrs.le.0.0) dorrsw = .false.") with cloud bucketing: the list holds first the
// cloud-free sunlit columns, then the sunlit columns with cloud (cldfra > 0 in some layer), each in tile order.  Columns are
// independent, so the order changes no result; it puts columns that take the cloudy-layer path of the solver (two extra
// two-stream evaluations per cloudy layer) into the same warps instead of idling 31 lanes for one.  Deterministic: one block
// scans a 1024-column segment, segment offsets come from a first counting pass.
__device__ __forceinline__ void sunlit_class(const Geo &G, const float *__restrict__ xcoszen, const float *__restrict__ cldfra3d, int tc,
                                             bool &sun, bool &cloudy) {
  sun = false; cloudy = false;
  if (tc >= G.ncol_tile) return;
  int i, j; G.ij(tc, i, j);
  sun = xcoszen ? !(xcoszen[G.at2(i, j)] <= 0.0f) : true;     // no xcoszen: every column (the LW list)
  if (sun && cldfra3d)
    for (int k = G.kts; k <= G.kte; k++) cloudy = cloudy || cldfra3d[G.at3(i, k, j)] > 0.f;
}
__global__ void k_count_sunlit(Geo G, const float *__restrict__ xcoszen, const float *__restrict__ cldfra3d, int nseg, int *__restrict__ segcount) {
  const int tc = blockIdx.x * blockDim.x + threadIdx.x;
  bool sun, cloudy;
  sunlit_class(G, xcoszen, cldfra3d, tc, sun, cloudy);
  const int n0 = __syncthreads_count(sun && !cloudy), n1 = __syncthreads_count(sun && cloudy);
  if (threadIdx.x == 0) { segcount[blockIdx.x] = n0; segcount[nseg + blockIdx.x] = n1; }
}
__global__ void k_scan_segments(int nseg, int *__restrict__ segcount, int *__restrict__ total) {
  // single thread block; nseg is small (2 * ncol/1024)
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nseg; base += blockDim.x) {
    int idx = base + threadIdx.x;
    int v = idx < nseg ? segcount[idx] : 0;
    // inclusive scan in shared memory (Hillis-Steele)
    __shared__ int buf[1024];
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < blockDim.x; off <<= 1) {
      int t = threadIdx.x >= off ? buf[threadIdx.x - off] : 0;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    if (idx < nseg) segcount[idx] = carry + buf[threadIdx.x] - v;   // exclusive offset
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry += buf[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
__global__ void k_fill_sunlit(Geo G, const float *__restrict__ xcoszen, const float *__restrict__ cldfra3d, int nseg, const int *__restrict__ segoff,
                              int *__restrict__ cols) {
  const int tc = blockIdx.x * blockDim.x + threadIdx.x;
  bool sun, cloudy;
  sunlit_class(G, xcoszen, cldfra3d, tc, sun, cloudy);
  __shared__ int wcount[2][32];
  const unsigned b0 = __ballot_sync(0xffffffffu, sun && !cloudy), b1 = __ballot_sync(0xffffffffu, sun && cloudy);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { wcount[0][w] = __popc(b0); wcount[1][w] = __popc(b1); }
  __syncthreads();
  if (!sun) return;
  const int q = cloudy ? 1 : 0;
  int off = segoff[q * nseg + blockIdx.x];
  for (int v = 0; v < w; v++) off += wcount[q][v];
  cols[off + __popc((q ? b1 : b0) & ((1u << lane) - 1u))] = tc;
}

// `slot` selects one of two scratch buffers: the SW list (0) and the LW list (1) of a chained radiation step are built on
// different streams at the same time
static int *g_seg[2] = {nullptr, nullptr}; static int g_seg_cap[2] = {0, 0};
void launch_compact_sunlit(const Geo &g, const float *xcoszen, const float *cldfra3d, int *cols, int *count, int slot, cudaStream_t s) {
  int nseg = (g.ncol_tile + 1023) / 1024;
  if (2 * nseg > g_seg_cap[slot]) { if (g_seg[slot]) cudaFree(g_seg[slot]); cudaMalloc(&g_seg[slot], sizeof(int) * 2 * nseg); g_seg_cap[slot] = 2 * nseg; }
  k_count_sunlit<<<nseg, 1024, 0, s>>>(g, xcoszen, cldfra3d, nseg, g_seg[slot]);
  k_scan_segments<<<1, 1024, 0, s>>>(2 * nseg, g_seg[slot], count);
  k_fill_sunlit<<<nseg, 1024, 0, s>>>(g, xcoszen, cldfra3d, nseg, g_seg[slot], cols);
  count_launch(3);
}

// Night columns and coszr: SW:10332, 11173-11199.  One thread per tile column.
__global__ void k_sw_night(SwArgs a) {
  int tc = blockIdx.x * blockDim.x + threadIdx.x;
  if (tc >= a.geo.ncol_tile) return;
  int i, j; a.geo.ij(tc, i, j);
  size_t ij = a.geo.at2(i, j);
  float cz = a.xcoszen[ij];
  a.coszr[ij] = cz;
  if (cz <= 0.0f) {
    if (a.swupt) {
      a.swupt[ij] = 0.f; a.swuptc[ij] = 0.f; a.swuptcln[ij] = 0.f; a.swdnt[ij] = 0.f; a.swdntc[ij] = 0.f; a.swdntcln[ij] = 0.f;
      a.swupb[ij] = 0.f; a.swupbc[ij] = 0.f; a.swupbcln[ij] = 0.f; a.swdnb[ij] = 0.f; a.swdnbc[ij] = 0.f; a.swdnbcln[ij] = 0.f;
      a.swvisdir[ij] = 0.f; a.swvisdif[ij] = 0.f; a.swnirdir[ij] = 0.f; a.swnirdif[ij] = 0.f;
    }
    if (a.swuptclnc) { a.swuptclnc[ij] = 0.f; a.swdntclnc[ij] = 0.f; a.swupbclnc[ij] = 0.f; a.swdnbclnc[ij] = 0.f; }
    a.swddir[ij] = 0.f; a.swddni[ij] = 0.f; a.swddif[ij] = 0.f; a.swcf[ij] = 0.f;
    if (a.dbg.laytrop) a.dbg.laytrop[tc] = -1;
  }
}
void launch_sw_night(const SwArgs &a, cudaStream_t s) {
  k_sw_night<<<(a.geo.ncol_tile + 255) / 256, 256, 0, s>>>(a);
  count_launch();
}

// ------------------------------------------------------------------------------------------------------
// McICA masks: generate_stochastic_clouds(_sw) with icld = 2 (maximum-random), irng = 0 (kissvec).
//
// The reference draws one serial KISS chain per column (sub-column major, layer minor: SW:1745-1750), 112*nlay or
// 140*nlay steps.  Here every (column, sub-column) thread JUMPS to its own position in that chain and then draws its nlay
// numbers, which is exact because each of the four KISS sub-generators has a closed-form k-step map:
//   s1  LCG x -> 69069 x + 1327217885 (mod 2^32):        x_k = A_k x + C_k (mod 2^32)
//   s2  xorshift (13, -17, 5), linear over GF(2):        x_k = M_k x, M_k stored as the 32 images of the unit vectors
//   s3  multiply-with-carry x -> 18000 (x & 65535) + (x >> 16): with m = 18000*2^16 - 1, 18000*2^16 = 1 (mod m) gives
//       x_{n+1} = 18000 x_n (mod m) as integers in [0, m), so x_k = x_0 * 18000^k mod m    (seeds are < 1e9 < m)
//   s4  the same with 30903.
// The per-sub-column constants (A, C, M[32], J3, J4) are built on the host for k = permuteseed + g*nlay.
// The maximum-random rescale (SW:1762-1772) only couples layers of the same sub-column, so it is applied on the fly.
struct KissJump { uint32_t A, C, J3, J4, M[32]; };
static KissJump *g_jump[2] = {nullptr, nullptr};
static int g_jump_key[2][3] = {{0, 0, 0}, {0, 0, 0}};

__global__ void __launch_bounds__(256) k_mcica(McicaArgs a, const KissJump *__restrict__ jump) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int g = blockIdx.y;
  if (c >= a.ncols) return;
  const int tc = a.cols ? a.cols[c] : a.col0 + c;
  int i, j; a.geo.ij(tc, i, j);
  const int kts = a.geo.kts;
  Kiss K;
  {
    float pm[4];
    for (int l = 0; l < 4; l++) {
      float play = a.p3d[a.geo.at3(i, kts + l, j)] / 100.f;       // p1d = p3d/100 ; play = p1d
      pm[l] = play * 1.e2f;                                        // pmid = play*1.e2
    }
    const uint32_t s1 = (uint32_t)(int32_t)((pm[0] - (float)(int)pm[0]) * 1000000000.f);
    const uint32_t s2 = (uint32_t)(int32_t)((pm[1] - (float)(int)pm[1]) * 1000000000.f);
    const uint32_t s3 = (uint32_t)(int32_t)((pm[2] - (float)(int)pm[2]) * 1000000000.f);
    const uint32_t s4 = (uint32_t)(int32_t)((pm[3] - (float)(int)pm[3]) * 1000000000.f);
    const KissJump &J = jump[g];
    K.s1 = J.A * s1 + J.C;
    uint32_t x = 0u;
#pragma unroll
    for (int b = 0; b < 32; b++) x ^= ((s2 >> b) & 1u) ? J.M[b] : 0u;
    K.s2 = x;
    K.s3 = (uint32_t)(((unsigned long long)s3 * J.J3) % 1179647999ull);     // 18000 * 65536 - 1
    K.s4 = (uint32_t)(((unsigned long long)s4 * J.J4) % 2025259007ull);     // 30903 * 65536 - 1
  }
  float prev = 0.f, omc_prev = 1.f;
  uint32_t word = 0u;
  const bool havecf = a.icloud != 0 && a.cldfra3d;
  const float *pcf = havecf ? a.cldfra3d + a.geo.at3(i, kts, j) : nullptr;      // level stride = ni
  const int kstride = a.geo.ni;
  for (int l = 0; l < a.nlay; l++) {
    float cf = 0.f;
    if (l < a.nz && havecf) { cf = *pcf; pcf += kstride; }
    if (cf < 1.0e-20f) cf = 0.f;
    const float omc = 1.0f - cf;
    float x = K.next();
    if (l > 0) {
      if (prev > omc_prev) x = prev; else x = x * omc_prev;
    }
    prev = x; omc_prev = omc;
    if (x >= omc) word |= 1u << (l & 31);
    if ((l & 31) == 31 || l == a.nlay - 1) {
      a.mask[((size_t)g * a.W + (l >> 5)) * a.cap + c] = word;
      if (word) atomicOr(&a.anyc[(size_t)(l >> 5) * a.cap + c], word);
      word = 0u;
    }
  }
}

static void build_jump_table(int which, int nlay, int ngpt, int permuteseed) {
  if (g_jump[which] && g_jump_key[which][0] == nlay && g_jump_key[which][1] == ngpt && g_jump_key[which][2] == permuteseed) return;
  std::vector<KissJump> T(ngpt);
  const unsigned long long m3 = 1179647999ull, m4 = 2025259007ull;
  // transform after k steps, advanced incrementally: start with k = permuteseed, then + nlay per sub-column
  uint32_t A = 1u, C = 0u, M[32];
  unsigned long long J3 = 1ull, J4 = 1ull;
  for (int b = 0; b < 32; b++) M[b] = 1u << b;
  auto advance = [&](int steps) {
    for (int q = 0; q < steps; q++) {
      A = 69069u * A; C = 69069u * C + 1327217885u;
      for (int b = 0; b < 32; b++) { uint32_t x = M[b]; x ^= x << 13; x ^= x >> 17; x ^= x << 5; M[b] = x; }
      J3 = (J3 * 18000ull) % m3; J4 = (J4 * 30903ull) % m4;
    }
  };
  advance(permuteseed);
  for (int gq = 0; gq < ngpt; gq++) {
    T[gq].A = A; T[gq].C = C; T[gq].J3 = (uint32_t)J3; T[gq].J4 = (uint32_t)J4;
    for (int b = 0; b < 32; b++) T[gq].M[b] = M[b];
    advance(nlay);
  }
  if (!g_jump[which]) cudaMalloc(&g_jump[which], sizeof(KissJump) * 160);
  cudaMemcpy(g_jump[which], T.data(), sizeof(KissJump) * ngpt, cudaMemcpyHostToDevice);
  g_jump_key[which][0] = nlay; g_jump_key[which][1] = ngpt; g_jump_key[which][2] = permuteseed;
}

void launch_mcica(const McicaArgs &a, cudaStream_t s) {
  const int which = a.lw_buffer ? 1 : 0;
  build_jump_table(which, a.nlay, a.ngpt, a.permuteseed);
  cudaMemsetAsync(a.anyc, 0, sizeof(uint32_t) * (size_t)a.W * a.cap, s);
  dim3 grid((a.ncols + 255) / 256, a.ngpt);
  k_mcica<<<grid, 256, 0, s>>>(a, g_jump[which]);
  count_launch();
}

// ------------------------------------------------------------------------------------------------------
// Cloud optical properties of one (layer, band): cldprmc_sw SW:2100-2390, inflag >= 2 path.
// Returns an ARC_ERR_* code.  Inputs are the in-cloud paths / sizes of the layer.
__device__ inline int sw_cloud_optics(const DevTables &tb, int ib /*0..13*/, int iceflag, int liqflag, float ciwp, float clwp,
                                      float cswp, float radice, float radliq, float radsno, float &taucmc, float &ssacmc,
                                      float &asmcmc, float &taormc) {
  const float cldmin = 1.e-20f;
  float extcoice = 0.f, ssacoice = 0.f, gice = 0.f, forwice = 0.f;
  float extcosno = 0.f, ssacosno = 0.f, gsno = 0.f, forwsno = 0.f;
  float extcoliq = 0.f, ssacoliq = 0.f, gliq = 0.f, forwliq = 0.f;
  if ((ciwp + cswp) == 0.0f) {
  } else if (iceflag >= 3) {
    if (radice < 5.0f || radice > 140.0f) return ARC_ERR_RADIUS;
    float factor = (radice - 2.f) / 3.f;
    int index = (int)factor;
    if (index == 46) index = 45;
    float fint = factor - (float)index;
    const int o = (index - 1) + 46 * ib;
    extcoice = tb.sw_extice3[o] + fint * (tb.sw_extice3[o + 1] - tb.sw_extice3[o]);
    ssacoice = tb.sw_ssaice3[o] + fint * (tb.sw_ssaice3[o + 1] - tb.sw_ssaice3[o]);
    gice = tb.sw_asyice3[o] + fint * (tb.sw_asyice3[o + 1] - tb.sw_asyice3[o]);
    float fdelta = tb.sw_fdlice3[o] + fint * (tb.sw_fdlice3[o + 1] - tb.sw_fdlice3[o]);
    if (fdelta < 0.0f || fdelta > 1.0f) return ARC_ERR_RADIUS;
    forwice = fdelta + 0.5f / ssacoice;
    if (forwice > gice) forwice = gice;
  } else {
    return ARC_ERR_UNSUPPORTED;
  }
  if (cswp > 0.0f && iceflag == 5) {
    if (radsno < 5.0f || radsno > 140.0f) return ARC_ERR_RADIUS;
    float factor = (radsno - 2.f) / 3.f;
    int index = (int)factor;
    if (index == 46) index = 45;
    float fint = factor - (float)index;
    const int o = (index - 1) + 46 * ib;
    extcosno = tb.sw_extice3[o] + fint * (tb.sw_extice3[o + 1] - tb.sw_extice3[o]);
    ssacosno = tb.sw_ssaice3[o] + fint * (tb.sw_ssaice3[o + 1] - tb.sw_ssaice3[o]);
    gsno = tb.sw_asyice3[o] + fint * (tb.sw_asyice3[o + 1] - tb.sw_asyice3[o]);
    float fdelta = tb.sw_fdlice3[o] + fint * (tb.sw_fdlice3[o + 1] - tb.sw_fdlice3[o]);
    if (fdelta < 0.0f || fdelta > 1.0f) return ARC_ERR_RADIUS;
    forwsno = fdelta + 0.5f / ssacosno;
    if (forwsno > gsno) forwsno = gsno;
  }
  if (clwp == 0.0f) {
  } else if (liqflag == 1) {
    if (radliq < 1.5f || radliq > 60.f) return ARC_ERR_RADIUS;
    int index = (int)(radliq - 1.5f);
    if (index == 0) index = 1;
    if (index == 58) index = 57;
    float fint = radliq - 1.5f - (float)index;
    const int o = (index - 1) + 58 * ib;
    extcoliq = tb.sw_extliq1[o] + fint * (tb.sw_extliq1[o + 1] - tb.sw_extliq1[o]);
    ssacoliq = tb.sw_ssaliq1[o] + fint * (tb.sw_ssaliq1[o + 1] - tb.sw_ssaliq1[o]);
    if (fint < 0.f && ssacoliq > 1.f) ssacoliq = tb.sw_ssaliq1[o];
    gliq = tb.sw_asyliq1[o] + fint * (tb.sw_asyliq1[o + 1] - tb.sw_asyliq1[o]);
    forwliq = gliq * gliq;
  }
  float tauliqorig = clwp * extcoliq;
  float tauiceorig = ciwp * extcoice;
  float ssaliq = ssacoliq * (1.f - forwliq) / (1.f - forwliq * ssacoliq);
  float tauliq = (1.f - forwliq * ssacoliq) * tauliqorig;
  float ssaice = ssacoice * (1.f - forwice) / (1.f - forwice * ssacoice);
  float tauice = (1.f - forwice * ssacoice) * tauiceorig;
  float scatliq = ssaliq * tauliq, scatice = ssaice * tauice, scatsno;
  if (iceflag < 5) {
    taormc = tauliqorig + tauiceorig;
    scatsno = 0.0f;
    taucmc = tauliq + tauice;
  } else {
    float tausnoorig = cswp * extcosno;
    taormc = tauliqorig + tauiceorig + tausnoorig;
    float ssasno = ssacosno * (1.f - forwsno) / (1.f - forwsno * ssacosno);
    float tausno = (1.f - forwsno * ssacosno) * tausnoorig;
    scatsno = ssasno * tausno;
    taucmc = tauliq + tauice + tausno;
  }
  if (taucmc == 0.f) taucmc = cldmin;
  if (scatice == 0.f) scatice = cldmin;
  if (scatsno == 0.f) scatsno = cldmin;
  if (iceflag < 5) ssacmc = (scatliq + scatice) / taucmc;
  else ssacmc = (scatliq + scatice + scatsno) / taucmc;
  if (iceflag == 3 || iceflag == 4) {
    asmcmc = (1.0f / (scatliq + scatice)) *
             (scatliq * (gliq - forwliq) / (1.0f - forwliq) + scatice * ((gice - forwice) / (1.0f - forwice)));
  } else {
    asmcmc = (1.0f / (scatliq + scatice + scatsno)) *
             (scatliq * (gliq - forwliq) / (1.0f - forwliq) + scatice * ((gice - forwice) / (1.0f - forwice)) +
              scatsno * ((gsno - forwsno) / (1.0f - forwsno)));
  }
  return 0;
}

// cldprmc LW:2653-2914 for one (layer, band)
__device__ inline int lw_cloud_optics(const DevTables &tb, int ib /*0..15*/, int iceflag, int liqflag, float ciwp, float clwp,
                                      float cswp, float radice, float radliq, float radsno, float &taucmc) {
  float abscoice = 0.f, abscosno = 0.f, abscoliq = 0.f;
  if ((ciwp + cswp) == 0.0f) {
  } else if (iceflag >= 3) {
    if (radice < 5.0f || radice > 140.0f) return ARC_ERR_RADIUS;
    float factor = (radice - 2.f) / 3.f;
    int index = (int)factor;
    if (index == 46) index = 45;
    float fint = factor - (float)index;
    const int o = (index - 1) + 46 * ib;
    abscoice = tb.lw_absice3[o] + fint * (tb.lw_absice3[o + 1] - (tb.lw_absice3[o]));
  } else {
    return ARC_ERR_UNSUPPORTED;
  }
  if (cswp > 0.0f && iceflag == 5) {
    if (radsno < 5.0f || radsno > 140.0f) return ARC_ERR_RADIUS;
    float factor = (radsno - 2.f) / 3.f;
    int index = (int)factor;
    if (index == 46) index = 45;
    float fint = factor - (float)index;
    const int o = (index - 1) + 46 * ib;
    abscosno = tb.lw_absice3[o] + fint * (tb.lw_absice3[o + 1] - (tb.lw_absice3[o]));
  }
  if (clwp == 0.0f) {
  } else if (liqflag == 1) {
    if (radliq < 2.5f || radliq > 60.f) return ARC_ERR_RADIUS;
    int index = (int)(radliq - 1.5f);
    if (index == 0) index = 1;
    if (index == 58) index = 57;
    float fint = radliq - 1.5f - (float)index;
    const int o = (index - 1) + 58 * ib;
    abscoliq = tb.lw_absliq1[o] + fint * (tb.lw_absliq1[o + 1] - (tb.lw_absliq1[o]));
  }
  taucmc = ciwp * abscoice + clwp * abscoliq + cswp * abscosno;
  return 0;
}

// Self-test tap: the pressure / temperature indices of setcoef for arbitrary (p, T) pairs, through the same pt_coef code
// the prep kernels use (jp | jt << 8 | jt1 << 12).
__global__ void k_selftest_pt(DevTables tb, const float *__restrict__ p, const float *__restrict__ t, int n, int *__restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  PTCoef c;
  pt_coef(tb.sw_preflog, tb.sw_tref, p[q], t[q], c);
  out[q] = c.jp | (c.jt << 8) | (c.jt1 << 12);
}
void launch_selftest_pt(const DevTables &tb, const float *p, const float *t, int n, int *packed, cudaStream_t s) {
  k_selftest_pt<<<(n + 255) / 256, 256, 0, s>>>(tb, p, t, n, packed);
  count_launch();
}

// ------------------------------------------------------------------------------------------------------
// SW column preparation.  One thread per sunlit column.
constexpr int PREP_MAXLAY = 160;

__global__ void __launch_bounds__(128) k_sw_prep(SwArgs a) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.ncols) return;
  const Geo &G = a.geo;
  const DevTables &tb = a.tb;
  const SwWs &ws = a.ws;
  const int tc = ws.cols[c];
  int i, j; G.ij(tc, i, j);
  const size_t ij = G.at2(i, j);
  const int kts = G.kts, nz = G.kte - G.kts + 1, nlay = nz + 1;
  const size_t cap = ws.cap;
  const float amdw = 1.607793f, amdo = 0.603461f;
  const float co2 = 379.e-6f, ch4 = 1774.e-9f, n2o = 319.e-9f, o2 = 0.209488f;
  (void)n2o;
  int inflg, iceflg, liqflg;
  cloud_flags(a.cf, inflg, iceflg, liqflg);

  auto COEF = [&](int f, int l) -> float & { return ws.coef[coef_index(f, l, c, cap, SWC_N)]; };
  auto AER = [&](int b, int q, int l) -> float & { return ws.aer[(((size_t)b * 3 + q) * nlay + l) * cap + c]; };
  auto CLD = [&](int b, int q, int l) -> float & { return ws.cld[(((size_t)b * 4 + q) * nlay + l) * cap + c]; };

  // column scalars
  const float coszrs = a.xcoszen[ij];
  {
    float cossza = coszrs;
    if (cossza <= 1.e-10f) cossza = 1.e-10f;
    float asdir, asdif, aldir, aldif;
    if (a.sf_surface_physics == 8 && a.cf.xland[ij] < 1.5f) {
      asdir = a.alswvisdir[ij]; asdif = a.alswvisdif[ij]; aldir = a.alswnirdir[ij]; aldif = a.alswnirdif[ij];
    } else { asdir = asdif = aldir = aldif = a.albedo[ij]; }
    ws.colf[(size_t)SWF_MU0 * cap + c] = cossza;
    ws.colf[(size_t)SWF_ALBDIR_NIR * cap + c] = aldir;
    ws.colf[(size_t)SWF_ALBDIF_NIR * cap + c] = aldif;
    ws.colf[(size_t)SWF_ALBDIR_UV * cap + c] = asdir;
    ws.colf[(size_t)SWF_ALBDIF_UV * cap + c] = asdif;
    // inatm_sw: adjflx = adjes (=1, dyofyr = 0); adjflux(ib) = adjflx * scon/1368.22  (SW:9746-9762)
    const float solvar = a.solcon / 1.36822e+03f;
    ws.colf[(size_t)SWF_ADJFLUX * cap + c] = 1.0f * solvar;
  }

  // log2 of the 14 Angstrom bases 0.4 / wavemid (the level-independent half of the reference's ** , SW:11008)
  double l2wave[NBSW];
#pragma unroll
  for (int b = 0; b < NBSW; b++) l2wave[b] = glm::powf_log2(0.4f / tb.wavemid[b]);

  unsigned char jps[PREP_MAXLAY];
  float o3top_ref = 0.f, o31d_top = 0.f;   // o3mmr(nz), o31d(nz) for the shifted climatology above the model top
  float h2o_prev = 0.f;
  int laytrop = 0;
  float aersum[NBSW];
  for (int b = 0; b < NBSW; b++) aersum[b] = 0.f;
  const float thresh = 1.e-9f;
  const float stpfac = 296.f / 1013.f;
  int err = 0;

  for (int l = 0; l < nlay; l++) {
    const bool model = l < nz;
    const int k = kts + l;
    // ---- pressures / temperature of the layer (hPa, K)
    float plev_b, plev_t, play, tlay;
    if (model) {
      plev_b = a.p8w[G.at3(i, k, j)] / 100.f;
      plev_t = a.p8w[G.at3(i, k + 1, j)] / 100.f;
      play = a.p3d[G.at3(i, k, j)] / 100.f;
      tlay = a.cf.t3d[G.at3(i, k, j)];
    } else {
      plev_b = a.p8w[G.at3(i, k, j)] / 100.f;     // plev(kte+1)
      plev_t = 1.0e-5f;                            // plev(kte+2)
      play = 0.5f * plev_b;
      tlay = a.t8w[G.at3(i, k, j)] + 0.0f;        // tlev(kte+1) + 0
    }
    const float pdel = plev_b - plev_t;
    // ---- water vapour, ozone (vmr)
    float h2ovmr;
    if (model) { h2ovmr = layer_qv(a.cf, G.at3(i, k, j)) * amdw; h2o_prev = h2ovmr; }
    else h2ovmr = h2o_prev;
    const float o3mmr = o3_clim(tb, plev_b, plev_t);
    float o3vmr = o3mmr * amdo;
    if (a.o33d && a.o3input == 2) {
      if (model) { o3vmr = a.o33d[G.at3(i, k, j)]; if (l == nz - 1) { o31d_top = o3vmr; o3top_ref = o3mmr; } }
      else {
        o3vmr = o31d_top - o3top_ref * amdo + o3mmr * amdo;
        if (o3vmr <= 0.f) o3vmr = o3mmr * amdo;
      }
    }
    // ---- inatm_sw: column amounts
    const float coldry = coldry_of(plev_b, plev_t, h2ovmr);
    const float wk_h2o = coldry * h2ovmr, wk_co2 = coldry * co2, wk_o3 = coldry * o3vmr, wk_ch4 = coldry * ch4,
                wk_o2 = coldry * o2;
    // ---- setcoef_sw
    PTCoef pc;
    pt_coef(tb.sw_preflog, tb.sw_tref, play, tlay, pc);
    const float water = wk_h2o / coldry;
    const float scalefac = play * stpfac / tlay;
    float forfac, forfrac, selffac, selffrac;
    int indfor, indself;
    if (!(pc.plog <= 4.56f)) {
      laytrop = laytrop + 1;
      forfac = scalefac / (1.f + water);
      float factor = (332.0f - tlay) / 36.0f;
      indfor = min(2, max(1, (int)factor));
      forfrac = factor - (float)indfor;
      selffac = water * forfac;
      factor = (tlay - 188.0f) / 7.2f;
      indself = min(9, max(1, (int)factor - 7));
      selffrac = factor - (float)(indself + 7);
    } else {
      forfac = scalefac / (1.f + water);
      float factor = (tlay - 188.0f) / 36.0f;
      indfor = 3;
      forfrac = factor - 1.0f;
      selffac = 0.f; selffrac = 0.f; indself = 0;
    }
    float colh2o = 1.e-20f * wk_h2o, colco2 = 1.e-20f * wk_co2, colo3 = 1.e-20f * wk_o3, colch4 = 1.e-20f * wk_ch4,
          colo2 = 1.e-20f * wk_o2;
    const float colmol = 1.e-20f * coldry + colh2o;
    if (colco2 == 0.f) colco2 = 1.e-32f * coldry;
    if (colch4 == 0.f) colch4 = 1.e-32f * coldry;
    if (colo2 == 0.f) colo2 = 1.e-32f * coldry;
    COEF(SWC_FAC00, l) = pc.fac00; COEF(SWC_FAC01, l) = pc.fac01; COEF(SWC_FAC10, l) = pc.fac10; COEF(SWC_FAC11, l) = pc.fac11;
    COEF(SWC_H2O, l) = colh2o; COEF(SWC_CO2, l) = colco2; COEF(SWC_O3, l) = colo3; COEF(SWC_CH4, l) = colch4;
    COEF(SWC_O2, l) = colo2; COEF(SWC_MOL, l) = colmol;
    COEF(SWC_SELFFAC, l) = selffac; COEF(SWC_SELFFRAC, l) = selffrac; COEF(SWC_FORFAC, l) = forfac; COEF(SWC_FORFRAC, l) = forfrac;
    COEF(SWC_IDX, l) = __int_as_float(pack_idx(pc.jp, pc.jt, pc.jt1, indself, indfor, 0));
    jps[l] = (unsigned char)pc.jp;
    if (a.dbg.jp) {
      const size_t q = (size_t)tc * nlay + l;
      a.dbg.jp[q] = pc.jp; a.dbg.jt[q] = pc.jt; a.dbg.jt1[q] = pc.jt1; a.dbg.indfor[q] = indfor; a.dbg.indself[q] = indself;
      a.dbg.fac00[q] = pc.fac00; a.dbg.fac01[q] = pc.fac01; a.dbg.fac10[q] = pc.fac10; a.dbg.fac11[q] = pc.fac11;
    }
    // ---- aerosol band optics (SW:10967-11030)
    float t300 = 0.f, t999 = 0.f, t400 = 0.f, w4 = 0.f, w6 = 0.f, g4 = 0.f, g6 = 0.f, lograt = 0.f;
    bool chem = false;
    if (model && a.aer_ra_feedback == 1) {
      const size_t q = G.at3(i, k, j);
      t300 = a.tauaer300[q]; t999 = a.tauaer999[q];
      if (t300 > thresh && t999 > thresh) {
        chem = true;
        t400 = a.tauaer400[q]; w4 = a.waer400[q]; w6 = a.waer600[q]; g4 = a.gaer400[q]; g6 = a.gaer600[q];
        lograt = glm::logf_(t300 / t999) / glm::logf_(999.f / 300.f);
      }
    }
    for (int b = 0; b < NBSW; b++) {
      float taua = 0.f, ssaa = 1.f, asma = 0.f;
      if (model && a.tauaer3d_sw) {
        const size_t q4 = G.at3(i, k, j) + G.n3() * (size_t)b;
        taua = a.tauaer3d_sw[q4]; ssaa = a.ssaaer3d_sw[q4]; asma = a.asyaer3d_sw[q4];
      }
      if (chem) {
        const float wavemid = tb.wavemid[b];
        taua = t400 * glm::powf_with(0.4f / wavemid, lograt, l2wave[b]);
        float slope = (w6 - w4) / .2f;
        ssaa = slope * (wavemid - .6f) + w6;
        if (ssaa < 0.4f) ssaa = 0.4f;
        if (ssaa >= 1.0f) ssaa = 1.0f;
        slope = (g6 - g4) / .2f;
        asma = slope * (wavemid - .6f) + g6;
        if (asma < 0.5f) asma = 0.5f;
        if (asma >= 1.0f) asma = 1.0f;
      }
      if (model) aersum[b] = aersum[b] + taua;
      if (a.aerod) {
        // aer_opt = 1, iaer = 6: rrtmg_sw ignores taua / ssaa / asma and mixes the six ECMWF aerosol types from their optical
        // depths ecaer (SW:9313-9341; the adapter copies AEROD for the model layers and 0 for the extra top layer, SW:11083-11100)
        float zt = 0.f, zo = 0.f, za = 0.f;
        for (int ia = 0; ia < 6; ia++) {
          const float ec = model ? a.aerod[G.at3(i, k, j) + G.n3() * (size_t)ia] : 0.f;
          const float re = tb.sw_rsr[ia * 14 + b] * ec;
          zt = zt + re;
          zo = zo + re * tb.sw_rsr[84 + ia * 14 + b];
          za = za + re * tb.sw_rsr[84 + ia * 14 + b] * tb.sw_rsr[168 + ia * 14 + b];
        }
        if (zo != 0.f) za = za / zo;
        if (zt != 0.f) zo = zo / zt;
        taua = zt; ssaa = zo; asma = za;
      }
      AER(b, 0, l) = taua; AER(b, 1, l) = ssaa; AER(b, 2, l) = asma;
    }
    // ---- cloud physical properties and band optics (only where some sub-column is cloudy)
    const bool anycld = (ws.anyc[(size_t)(l >> 5) * cap + c] >> (l & 31)) & 1u;
    if (anycld) {
      LayerCloud lc;
      if (model) layer_cloud(a.cf, G, tb, i, j, k, tlay, tlay, pdel, inflg, iceflg, lc);
      else { lc.clwp = 0.f; lc.ciwp = 0.f; lc.cswp = 0.f; lc.rel = 10.f; lc.rei = 10.f; lc.res = 10.f; }
      const float cswp = iceflg == 5 ? lc.cswp : 0.f;        // inatm_sw copies cswp only for iceflag 5 (SW:9852)
      const float resn = iceflg == 5 ? lc.res : 0.f;
      const float cwp = lc.ciwp + lc.clwp + cswp;
      for (int b = 0; b < NBSW; b++) {
        float tc_ = 0.f, sc = 1.f, ac = 0.f, to = 0.f;
        if (cwp >= 1.e-20f) {
          int rc = sw_cloud_optics(tb, b, iceflg, liqflg, lc.ciwp, lc.clwp, cswp, lc.rei, lc.rel, resn, tc_, sc, ac, to);
          if (rc && !err) err = rc;
        }
        CLD(b, 0, l) = tc_; CLD(b, 1, l) = sc; CLD(b, 2, l) = ac; CLD(b, 3, l) = to;
      }
    }
  }
  // aerosol column checks (SW:11031-11047)
  if (a.aer_ra_feedback == 1) {
    for (int b = 0; b < NBSW; b++) {
      const float slope = aersum[b];
      if (slope < 0.f) { if (!err) err = ARC_ERR_NEG_AOD; }
      else if (slope > 6.f && !a.aerod) {            // (with iaer = 6 the capped array is the one rrtmg_sw ignores)
        atomicAdd(a.status + 1, 1);                   // the reference's "WARNING: Large total sw optical depth" (SW:11049-11066)
        for (int l = 0; l < nz; l++) AER(b, 0, l) = AER(b, 0, l) * 6.0f / slope;
      }
    }
  }
  ws.laytrop[c] = laytrop;
  if (a.dbg.laytrop) a.dbg.laytrop[tc] = laytrop;
  // layer where each band takes its solar source function (taumol16..29), emulating the reference loops literally
  for (int b = 0; b < NBSW; b++) {
    const int layreffr = c_sw[b].layreffr;
    const int band = b + 16;
    int last = -1;
    const bool upper = (band == 16 || band == 17 || band == 27 || band == 28 || band == 29);
    if (band == 26) {
      last = laytrop >= 1 ? laytrop - 1 : -1;       // lay == laysolfr == laytrop inside "lay <= laytrop"
    } else if (upper) {
      int laysolfr = nlay;                           // 1-based
      for (int lay = laytrop + 1; lay <= nlay; lay++) {
        if (lay >= 2 && jps[lay - 2] < layreffr && jps[lay - 1] >= layreffr) laysolfr = lay;
        if (lay == laysolfr) last = lay - 1;
      }
    } else {
      int laysolfr = laytrop;
      for (int lay = 1; lay <= laytrop; lay++) {
        if (lay < nlay && jps[lay - 1] < layreffr && jps[lay] >= layreffr) laysolfr = min(lay + 1, laytrop);
        if (lay == laysolfr) last = lay - 1;
      }
    }
    ws.laysol[(size_t)b * cap + c] = last;
  }
  set_status(a.status, err);
}
void launch_sw_prep(const SwArgs &a, cudaStream_t s) {
  k_sw_prep<<<(a.ncols + 127) / 128, 128, 0, s>>>(a);
  count_launch();
}

// ------------------------------------------------------------------------------------------------------
// cal_cldfra1 (module_radiation_driver.F:2886-3122; icloud = 1): Randall-1994 / Hong-1998 cloud fraction from the grid-scale
// condensate and the relative humidity with respect to a water / ice weighted saturation mixing ratio.  One thread per cell,
// i fastest.  EXP and ** are glibc's (glibc_math.cuh) and this file is compiled without FMA contraction: bit-exact with the
// reference's arithmetic.  f_q* : 1 = .TRUE., 0 = .FALSE., < 0 = not PRESENT.
struct CldfraArgs {
  Geo geo;
  const float *qv, *qc, *qi, *qs, *t_phy, *p_phy, *f_ice_phy;
  int f_qv, f_qc, f_qi, f_qs, mp_physics;
  float *cldfra; int *flag;
};
__global__ void __launch_bounds__(256) k_cal_cldfra1(CldfraArgs a) {
  const Geo &G = a.geo;
  const int nz = G.kte - G.kts + 1;
  const long n = (long)G.ncol_tile * nz;
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int ii = (int)(t % G.nci), k = G.kts + (int)((t / G.nci) % nz), j = G.jts + (int)(t / ((long)G.nci * nz));
  const size_t q = G.at3(G.its + ii, k, j);
  const float ALPHA0 = 100.f, GAMMA = 0.49f, QCLDMIN = 1.E-12f, PEXP = 0.25f, RHGRID = 1.0f;
  const float SVP1 = 0.61078f, SVP2 = 17.2693882f, SVPI2 = 21.8745584f, SVP3 = 35.86f, SVPI3 = 7.66f, SVPT0 = 273.15f;
  const float ep_2 = 287.f / 461.6f;
  const float tk = a.t_phy[q], pp = a.p_phy[q];
  const float tc = tk - SVPT0;
  const float esw = 1000.0f * SVP1 * glm::expf_(SVP2 * tc / (tk - SVP3));
  const float esi = 1000.0f * SVP1 * glm::expf_(SVPI2 * tc / (tk - SVPI3));
  const float qvsw = ep_2 * esw / (pp - esw);
  const float qvsi = ep_2 * esi / (pp - esi);
  float weight = 0.f, qcld = 0.f;
  const bool present = a.f_qi >= 0 && a.f_qc >= 0 && a.f_qs >= 0;
  if (present) {
    const bool fqi = a.f_qi > 0, fqc = a.f_qc > 0, fqs = a.f_qs > 0;
    const float qi = a.qi ? a.qi[q] : 0.f, qc = a.qc ? a.qc[q] : 0.f, qs = a.qs ? a.qs[q] : 0.f;
    if (fqi && fqc && fqs) { qcld = qi + qc + qs; weight = qcld < QCLDMIN ? 0.f : (qi + qs) / qcld; }
    if (fqi && fqc && !fqs) { qcld = qi + qc; weight = qcld < QCLDMIN ? 0.f : qi / qcld; }
    if (fqc && !fqi && !fqs) { qcld = qc; weight = qcld < QCLDMIN ? 0.f : (tk > 273.15f ? 0.f : 1.f); }
    if (fqc && !fqi && fqs && a.f_ice_phy) { qcld = qc + qs; weight = qcld < QCLDMIN ? 0.f : a.f_ice_phy[q]; }
    if (a.mp_physics == 5 || a.mp_physics == 15) {          // FER_MP_HIRES, FER_MP_HIRES_ADVECT (Registry.EM_COMMON)
      qcld = qc + qi;
      if (qcld < QCLDMIN) weight = 0.f; else { weight = qi / qcld; if (tc < -40.f) weight = 1.f; }
    }
  }
  const float qvs_weight = (1 - weight) * qvsw + weight * qvsi;
  float rhum = a.qv[q] / qvs_weight;
  float cf; int fl;
  if (!present || qcld < QCLDMIN) { cf = 0.f; fl = 1; }
  else if (rhum >= RHGRID) { cf = 1.f; fl = 2; }
  else {
    fl = 3;
    const float subsat = fmaxf(1.E-10f, RHGRID * qvs_weight - a.qv[q]);
    const float denom = glm::powf_(subsat, GAMMA);
    const float arg = fmaxf(-6.9f, -ALPHA0 * qcld / denom);
    rhum = fmaxf(1.E-10f, rhum);
    cf = glm::powf_(rhum / RHGRID, PEXP) * (1.f - glm::expf_(arg));
    if (cf < .01f) cf = 0.f;
  }
  a.cldfra[q] = cf;
  if (a.flag) a.flag[q] = fl;
}
void launch_cal_cldfra1(const Geo &G, const float *qv, const float *qc, const float *qi, const float *qs, int f_qv, int f_qc, int f_qi, int f_qs,
                        const float *t_phy, const float *p_phy, const float *f_ice_phy, int mp_physics, float *cldfra, int *flag, cudaStream_t s) {
  CldfraArgs a{G, qv, qc, qi, qs, t_phy, p_phy, f_ice_phy, f_qv, f_qc, f_qi, f_qs, mp_physics, cldfra, flag};
  const long n = (long)G.ncol_tile * (G.kte - G.kts + 1);
  k_cal_cldfra1<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a);
  count_launch();
}

// cal_cldfra2 (module_radiation_driver.F:2801-2874; icloud = 2): cloud fraction 1 where QC + QI (or QC alone) exceeds 1e-6, else 0.
__global__ void __launch_bounds__(256) k_cal_cldfra2(Geo G, const float *__restrict__ qc, const float *__restrict__ qi, int f_qc, int f_qi,
                                                     float *__restrict__ cldfra) {
  const int nz = G.kte - G.kts + 1;
  const long n = (long)G.ncol_tile * nz;
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int ii = (int)(t % G.nci), k = G.kts + (int)((t / G.nci) % nz), j = G.jts + (int)(t / ((long)G.nci * nz));
  const size_t q = G.at3(G.its + ii, k, j);
  const float thresh = 1.0e-6f;
  float cf = 0.f;
  if (f_qi && f_qc) cf = __fadd_rn(qc[q], qi[q]) > thresh ? 1.f : 0.f;
  else if (f_qc) cf = qc[q] > thresh ? 1.f : 0.f;
  cldfra[q] = cf;
}
void launch_cal_cldfra2(const Geo &G, const float *qc, const float *qi, int f_qc, int f_qi, float *cldfra, cudaStream_t s) {
  const long n = (long)G.ncol_tile * (G.kte - G.kts + 1);
  k_cal_cldfra2<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(G, qc, qi, f_qc, f_qi, cldfra);
  count_launch();
}

// cal_cldfra3 (module_radiation_driver.F:3140-3274; icloud = 3, G. Thompson's Sundqvist-type scheme) in two kernels.
//   k_cldfra3_cell    one thread per cell: first-guess cloud fraction from RH against a grid-size dependent threshold (land /
//                     ocean), the saturation mixing ratio qvsat it used, and theta of find_cloudLayers.
//   k_cldfra3_column  one thread per column: find_cloudLayers (DRV:3281-3468) with adjust_cloudIce / adjust_cloudH2O /
//                     adjust_cloudFinal (DRV:3472-3599), working in place on CLDFRA, qc, qi (INOUT in the reference) with the
//                     1-D work arrays theta, dz, qvsat in global scratch, lanes = neighbouring columns.
// rslf / rsif belong to module_mp_thompson, which is not in the reference repository: the published Flatau et al. (1992)
// polynomials as the Thompson scheme codes them (parity unpinned for these two functions).  Unfused arithmetic, glibc **.
__device__ __forceinline__ float thompson_rs(float P, float T, bool ice) {
  const float W[9] = {.611583699E03f, .444606896E02f, .143177157E01f, .264224321E-1f, .299291081E-3f, .203154182E-5f, .702620698E-8f,
                      .379534310E-11f, -.321582393E-13f};
  const float I[9] = {.609868993E03f, .499320233E02f, .184672631E01f, .402737184E-1f, .565392987E-3f, .521693933E-5f, .307839583E-7f,
                      .105785160E-9f, .161444444E-12f};
  const float X = fmaxf(-80.f, T - 273.16f);
  float es = ice ? I[8] : W[8];
#pragma unroll
  for (int n = 7; n >= 0; n--) es = (ice ? I[n] : W[n]) + X * es;
  es = fminf(es, P * 0.15f);
  return .622f * es / (P - es);
}
struct Cldfra3Args {
  Geo geo;
  const float *qv, *qs, *p, *t, *rho, *xland;
  float *qc, *qi, *cldfra;
  float *qvsat, *theta, *dz;          // scratch, (i,k,j) like the fields
  float rh_00l, rh_00o;
};
__global__ void __launch_bounds__(256) k_cldfra3_cell(Cldfra3Args a) {
  const Geo &G = a.geo;
  const int nz = G.kte - G.kts + 1;
  const long n = (long)G.ncol_tile * nz;
  const long tq = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tq >= n) return;
  const int ii = (int)(tq % G.nci), k = G.kts + (int)((tq / G.nci) % nz), j = G.jts + (int)(tq / ((long)G.nci * nz));
  const size_t q = G.at3(G.its + ii, k, j);
  const float TK = a.t[q], P = a.p[q], qv = a.qv[q];
  a.theta[q] = TK * glm::powf_(100000.0f / P, 287.05f / 1004.f);
  float cf = 0.0f, qvs;
  if (a.qc[q] > 1.E-6f || a.qi[q] >= 1.E-7f || a.qs[q] > 1.E-5f) { cf = 1.0f; qvs = qv; }
  else {
    const float TC = TK - 273.16f;
    const float qvsw = thompson_rs(P, TK, false), qvsi = thompson_rs(P, TK, true);
    if (TC >= -12.0f) qvs = qvsw;
    else if (TC < -20.0f) qvs = qvsi;
    else qvs = qvsw - (qvsw - qvsi) * (-12.0f - TC) / (-12.0f + 20.f);
    float RHUM = fmaxf(0.01f, fminf(qv / qvs, 0.9999f));
    const float RH_00 = (a.xland[G.at2(G.its + ii, j)] - 1.5f) > 0.f ? a.rh_00o : a.rh_00l;
    if (TC >= -12.0f) {
      RHUM = fminf(0.999f, RHUM);
      cf = fmaxf(0.0f, 1.0f - sqrtf((1.0f - RHUM) / (1.f - RH_00)));
    } else if (TC < -12.f && TC > -70.f && RHUM > a.rh_00l) {
      RHUM = fmaxf(0.01f, fminf(qv / qvs, 1.0f - 1.E-6f));
      cf = fmaxf(0.f, 1.0f - sqrtf((1.0f - RHUM) / (1.0f - a.rh_00l)));
    }
    cf = fminf(0.90f, cf);
  }
  a.cldfra[q] = cf;
  a.qvsat[q] = qvs;
}
__global__ void __launch_bounds__(128) k_cldfra3_column(Cldfra3Args a) {
  const Geo &G = a.geo;
  const int tc = blockIdx.x * blockDim.x + threadIdx.x;
  if (tc >= G.ncol_tile) return;
  int i, j; G.ij(tc, i, j);
  const int kts = G.kts, kte = G.kte;
  const size_t q0 = G.at3(i, kts, j), ks = (size_t)G.ni;          // level stride
  auto at = [&](int k) { return q0 + (size_t)(k - kts) * ks; };
#define CFR(k) a.cldfra[at(k)]
#define QC1(k) a.qc[at(k)]
#define QI1(k) a.qi[at(k)]
#define QVS(k) a.qvsat[at(k)]
#define T1(k) a.t[at(k)]
#define P1(k) a.p[at(k)]
#define R1(k) a.rho[at(k)]
#define TH(k) a.theta[at(k)]
#define DZ(k) a.dz[at(k)]
  const float entr = 0.5f;
  auto height = [&](int k) { return 44307.692f * (1.0f - glm::powf_(P1(k) / 101325.f, 0.190f)); };
  int k, k2, k_m12C = 0, k_m40C = 0, k_cldb, k_cldt, kbot;
  for (k = kte; k >= kts; k--) {
    if (T1(k) - 273.16f > -40.0f && P1(k) > 7000.0f) k_m40C = max(k_m40C, k);
    if (T1(k) - 273.16f > -12.0f && P1(k) > 10000.0f) k_m12C = max(k_m12C, k);
  }
  if (k_m40C <= kts) k_m40C = kts;
  if (k_m12C <= kts) k_m12C = kts;
  float Z2 = height(kte);
  for (k = kte - 1; k >= kts; k--) { const float Z1 = height(k); DZ(k + 1) = Z2 - Z1; Z2 = Z1; }
  DZ(kts) = DZ(kts + 1);
  for (k = kte - 3; k >= kts; k--) {      // tropopause: d(theta)/dz below 10 K per 1500 m over three levels, between 4 and 19 km
    const float ht1 = height(k), ht2 = height(k + 2);
    if ((((TH(k + 2) - TH(k)) / (ht2 - ht1)) < 10.f / 1500.f) && (ht1 < 19000.f) && (ht1 > 4000.f)) break;
  }
  const int k_tropo = max(kts + 2, k + 2);
  for (k = k_tropo + 1; k <= kte; k++) { const float c = CFR(k); if (c > 0.0f && c < 0.999f) CFR(k) = 0.f; }
  kbot = kts + 2;
  for (k = kbot; k <= k_m12C; k++)
    if ((TH(k) - TH(k - 1)) > 0.05E-3f * DZ(k)) break;
  kbot = max(kts + 1, k - 2);
  for (k = kts; k <= kbot; k++) { const float c = CFR(k); if (c > 0.0f && c < 0.999f) CFR(k) = 0.f; }
  auto cfr_at = [&](int kk) { return kk <= kte ? CFR(kk) : 0.f; };       // the Fortran does not bound k_m12C + 2 by kte
  // the two layer searches: ice clouds from the tropopause down to the -12 C level, water clouds from there to kbot
  for (int pass = 0; pass < 2; pass++) {
    const int klow = pass == 0 ? k_m12C : kbot;
    k_cldb = k_tropo;
    k = pass == 0 ? k_tropo : k_m12C + 2;
    while (k > klow) {
      k_cldt = 0;
      if (cfr_at(k) >= 0.01f) {
        k_cldt = k;
        for (k2 = k_cldt - 1; k2 >= klow; k2--)
          if (CFR(k2) < 0.01f || k2 == klow) { k_cldb = k2 + 1; break; }
      }
      if ((k_cldt - k_cldb + 1) >= 2) {
        // adjust_cloudIce / adjust_cloudH2O: an adiabatic-like condensate profile over the layer's depth, entrainment-reduced
        float tdz = 0.f;
        for (int kk = k_cldb; kk <= k_cldt; kk++) tdz = tdz + DZ(kk);
        const float max_wc = fabsf(QVS(k_cldt - 1) - QVS(k_cldb));
        float this_dz = 0.0f;
        for (int kk = k_cldb; kk <= k_cldt; kk++) {
          this_dz = kk == k_cldb ? this_dz + 0.5f * DZ(kk) : this_dz + DZ(kk);
          const float wc = fmaxf(1.E-6f, (max_wc * this_dz / tdz) * (1.f - entr));
          const float c = CFR(kk), T = T1(kk);
          if (pass == 0) {
            const float qi = QI1(kk);
            if (c > 0.01f && c < 0.99f && T >= 203.16f) QI1(kk) = qi + 0.1f * c * wc;
            else if (qi < 1.E-5f && c >= 0.99f && T >= 203.16f) QI1(kk) = qi + 0.01f * wc;
          } else {
            const float qc = QC1(kk);
            if (c > 0.01f && c < 0.99f && T < 298.16f && T >= 253.16f) QC1(kk) = qc + c * c * wc;
            else if (c >= 0.99f && qc < 1.E-5f && T < 298.16f && T >= 253.16f) QC1(kk) = qc + 0.1f * wc;
          }
        }
        k = k_cldb;
      } else {
        const float c = CFR(k_cldb);
        if (pass == 0) { if (c > 0.f && QI1(k_cldb) < 1.E-6f) QI1(k_cldb) = 1.E-5f * c; }
        else { if (c > 0.f && QC1(k_cldb) < 1.E-6f) QC1(k_cldb) = 1.E-5f * c; }
      }
      k = k - 1;
    }
  }
  // adjust_cloudFinal: more than 1.5 kg m-2 of made-up condensate below the tropopause is scaled back
  float lwp = 0.f, iwp = 0.f;
  for (k = kts; k <= k_tropo; k++)
    if (CFR(k) > 0.0f) { const float m = R1(k); lwp = lwp + QC1(k) * m * DZ(k); iwp = iwp + QI1(k) * m * DZ(k); }
  if (lwp > 1.5f) { const float xfac = 1.f / lwp; for (k = kts; k <= k_tropo; k++) { const float c = CFR(k); if (c > 0.01f && c < 0.99f) QC1(k) = QC1(k) * xfac; } }
  if (iwp > 1.5f) { const float xfac = 1.f / iwp; for (k = kts; k <= k_tropo; k++) { const float c = CFR(k); if (c > 0.01f && c < 0.99f) QI1(k) = QI1(k) * xfac; } }
#undef CFR
#undef QC1
#undef QI1
#undef QVS
#undef T1
#undef P1
#undef R1
#undef TH
#undef DZ
}
void launch_cal_cldfra3(const Geo &G, float *cldfra, const float *qv, float *qc, float *qi, const float *qs, const float *p, const float *t,
                        const float *rho, const float *xland, float gridkm, float *qvsat, float *theta, float *dz, cudaStream_t s) {
  Cldfra3Args a{G, qv, qs, p, t, rho, xland, qc, qi, cldfra, qvsat, theta, dz, 0.f, 0.f};
  a.rh_00l = 0.7f + sqrtf(1.f / (25.0f + gridkm * gridkm * gridkm));
  a.rh_00o = 0.81f + sqrtf(1.f / (50.0f + gridkm * gridkm * gridkm));
  const long n = (long)G.ncol_tile * (G.kte - G.kts + 1);
  k_cldfra3_cell<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a);
  k_cldfra3_column<<<(G.ncol_tile + 127) / 128, 128, 0, s>>>(a);
  count_launch(2);
}

// ozn_time_int (module_radiation_driver.F:3993-4098): ozmixt(i,k,j) = ozmixm(i,k,j,nm) * fact1 + ozmixm(i,k,j,np) * fact2 for the
// tile's (i, j) and all levsiz data levels; the month indices and weights come from the host (scalar date arithmetic).
__global__ void __launch_bounds__(256) k_ozn_time_int(Geo G, int levsiz, const float *__restrict__ m0, const float *__restrict__ m1, float fact1,
                                                      float fact2, float *__restrict__ ozmixt) {
  const long n = (long)G.ncol_tile * levsiz;
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int ii = (int)(t % G.nci), k = (int)((t / G.nci) % levsiz), j = G.jts + (int)(t / ((long)G.nci * levsiz));
  const size_t q = (size_t)(G.its + ii - G.ims) + (size_t)G.ni * ((size_t)k + (size_t)levsiz * (size_t)(j - G.jms));
  ozmixt[q] = __fadd_rn(__fmul_rn(m0[q], fact1), __fmul_rn(m1[q], fact2));
}
void launch_ozn_time_int(const Geo &G, int levsiz, const float *m0, const float *m1, float fact1, float fact2, float *ozmixt, cudaStream_t s) {
  const long n = (long)G.ncol_tile * levsiz;
  k_ozn_time_int<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(G, levsiz, m0, m1, fact1, fact2, ozmixt);
  count_launch();
}

// ozn_p_int (module_radiation_driver.F:4100-4234): the data-level ozone interpolated linearly in pressure to the model levels,
// top level first.  The reference walks a tile row with a shared starting level (kkstart = the smallest kupper of the row);
// because the model pressure rises monotonically towards the surface every column finds the same bracket it would find from
// its own kupper, so one thread per column carrying its own kupper reproduces it.  Above the data top the mixing ratio is
// scaled by p / pin(1), below the data bottom it is held.  Unfused arithmetic, IEEE division.
struct OznPin { float pin[ARC_OZN_MAXLEV]; };
// One kernel serves ozn_p_int and aer_p_int (DRV:4345-4506), which is the same interpolation per aerosol type with the model
// pressure in hPa (p * 0.01), the result multiplied by the layer's interface-pressure difference pf(k) - pf(k+1), and the column
// total over types and levels (types outer, levels inner, as the reference accumulates).
__global__ void __launch_bounds__(128) k_clim_p_int(Geo G, int levsiz, OznPin P, const float *__restrict__ p, float pscale, int nsrc,
                                                    const float *__restrict__ data, float *__restrict__ out, const float *__restrict__ pf,
                                                    float *__restrict__ total) {
  const int tc = blockIdx.x * blockDim.x + threadIdx.x;
  if (tc >= G.ncol_tile) return;
  int i, j; G.ij(tc, i, j);
  const size_t nlev = (size_t)G.ni * (size_t)levsiz * (size_t)(G.jme - G.jms + 1), n3 = G.n3();
  float tot = 0.f;
  for (int s = 0; s < nsrc; s++) {
    const float *__restrict__ dt = data + nlev * (size_t)s;
    float *__restrict__ o = out + n3 * (size_t)s;
    const size_t ob = (size_t)(i - G.ims) + (size_t)G.ni * (size_t)levsiz * (size_t)(j - G.jms);   // data(i, 1, j)
    int kupper = 1;                                                                                   // 1-based, as in the reference
    for (int k = G.kte; k >= G.kts; k--) {
      const size_t q = G.at3(i, k, j);
      const float pm = __fmul_rn(p[q], pscale);
      for (int kk = kupper; kk <= levsiz - 1; kk++)
        if (P.pin[kk - 1] < pm && pm <= P.pin[kk]) { kupper = kk; break; }
      float v;
      if (pm < P.pin[0]) v = div_rn(__fmul_rn(dt[ob], pm), P.pin[0]);
      else if (pm > P.pin[levsiz - 1]) v = dt[ob + (size_t)G.ni * (levsiz - 1)];
      else {
        const float dpu = __fsub_rn(pm, P.pin[kupper - 1]);
        const float dpl = __fsub_rn(P.pin[kupper], pm);
        v = div_rn(__fadd_rn(__fmul_rn(dt[ob + (size_t)G.ni * (kupper - 1)], dpl), __fmul_rn(dt[ob + (size_t)G.ni * kupper], dpu)),
                   __fadd_rn(dpl, dpu));
      }
      if (pf) v = __fmul_rn(v, __fsub_rn(pf[q], pf[G.at3(i, k + 1, j)]));
      o[q] = v;
    }
    if (total)
      for (int k = G.kts; k <= G.kte; k++) tot = __fadd_rn(tot, o[G.at3(i, k, j)]);
  }
  if (total) total[G.at2(i, j)] = tot;
}
void launch_clim_p_int(const Geo &G, int levsiz, const float *pin_host, const float *p, float pscale, int nsrc, const float *data, float *out,
                       const float *pf, float *total, cudaStream_t s) {
  OznPin P;
  for (int k = 0; k < ARC_OZN_MAXLEV; k++) P.pin[k] = k < levsiz ? pin_host[k] : 0.f;
  k_clim_p_int<<<(G.ncol_tile + 127) / 128, 128, 0, s>>>(G, levsiz, P, p, pscale, nsrc, data, out, pf, total);
  count_launch();
}
void launch_ozn_p_int(const Geo &G, int levsiz, const float *pin_host, const float *p, const float *ozmixt, float *o3vmr, cudaStream_t s) {
  launch_clim_p_int(G, levsiz, pin_host, p, 1.0f, 1, ozmixt, o3vmr, nullptr, nullptr, s);
}

// ------------------------------------------------------------------------------------------------------
// LW column preparation.  One thread per column (all columns; LW has no day/night gate).
__global__ void __launch_bounds__(128) k_lw_prep(LwArgs a) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.ncols) return;
  const Geo &G = a.geo;
  const DevTables &tb = a.tb;
  const LwWs &ws = a.ws;
  const int tc = a.ws.cols ? a.ws.cols[c] : a.col0 + c;
  int i, j; G.ij(tc, i, j);
  const size_t ij = G.at2(i, j);
  const int kts = G.kts, nz = G.kte - G.kts + 1, nlay = ws.nlay;
  const size_t cap = ws.cap;
  const float amdw = 1.607793f, amdo = 0.603461f, deltap = 4.f;
  const float co2 = 379.e-6f, ch4 = 1774.e-9f, n2o = 319.e-9f, o2 = 0.209488f;
  const float cfc11 = 0.251e-9f, cfc12 = 0.538e-9f, cfc22 = 0.169e-9f, ccl4 = 0.093e-9f;
  (void)cfc11; (void)cfc12; (void)cfc22; (void)ccl4;
  const float amd = 28.9660f, amw = 18.0160f;
  const float thresh = 1.e-9f;
  const float stpfac = 296.f / 1013.f;
  int inflg, iceflg, liqflg;
  cloud_flags(a.cf, inflg, iceflg, liqflg);

  auto COEF = [&](int f, int l) -> float & { return ws.coef[coef_index(f, l, c, cap, LWC_N)]; };

  // temperature profile interpolated to a pressure level (LW:12220-12243)
  auto varint = [&](float p) {
    const int nproflevs = 60;
    int klev = nproflevs;
    if (tb.pprof[nproflevs - 1] < p) {
      for (int LL = 2; LL <= nproflevs; LL++) { if (tb.pprof[LL - 1] < p) { klev = LL - 1; break; } }
    }
    float vark, vark1, wght;
    if (klev != nproflevs) {
      vark = tb.tprof[klev - 1]; vark1 = tb.tprof[klev];
      wght = (p - tb.pprof[klev - 1]) / (tb.pprof[klev] - tb.pprof[klev - 1]);
    } else { vark = tb.tprof[klev - 1]; vark1 = tb.tprof[klev - 1]; wght = 0.0f; }
    return wght * (vark1 - vark) + vark;
  };

  // interface pressures / temperatures.  plev(L), L = 1..nlay+1 ; tlev likewise (1-based as in the reference)
  // model part: plev(k) = p8w/100, tlev(k) = t8w, k = 1..nz+1.  Buffer: plev(L+1) = plev(L) - 4, plev(nlay+1) = 0.
  // tlev(L) = varint(L) + (tlev(nz) - varint(nz)) for L = nz+1..nlay+1 (note: overwrites tlev(nz+1));
  // tlay(L-1) = 0.5*(tlev(L) + tlev(L-1))  for the same L (overwrites tlay(nz)).
  const float tlev_nz = a.t8w[G.at3(i, kts + nz - 1, j)];
  const float plev_nz = a.p8w[G.at3(i, kts + nz - 1, j)] / 100.f;
  const float shift = tlev_nz - varint(plev_nz);

  const float tsfc = a.tsk[ij];
  ws.colf[(size_t)LWF_TBOUND * cap + c] = tsfc;
  ws.colf[(size_t)LWF_EMISS * cap + c] = a.emiss[ij];
  ws.colf[(size_t)LWF_TZ0 * cap + c] = a.t8w[G.at3(i, kts, j)];

  float plev_b = a.p8w[G.at3(i, kts, j)] / 100.f;       // plev(1)
  float tlev_b = a.t8w[G.at3(i, kts, j)];               // tlev(1)
  const float pz0 = plev_b;
  float h2o_top = 0.f, o3top_ref = 0.f, o31d_top = 0.f;
  float amttl = 0.f, wvttl = 0.f;
  int laytrop = 0;
  int err = 0;
  float aersum[NBLW];
  for (int b = 0; b < NBLW; b++) aersum[b] = 0.f;

  for (int l = 0; l < nlay; l++) {
    const int L = l + 1;                 // 1-based layer
    const bool model = L <= nz;
    const int k = kts + l;
    float plev_t, tlev_t, play, tlay;
    if (model) {
      plev_t = a.p8w[G.at3(i, k + 1, j)] / 100.f;
      play = a.p3d[G.at3(i, k, j)] / 100.f;
      tlay = a.cf.t3d[G.at3(i, k, j)];
      tlev_t = a.t8w[G.at3(i, k + 1, j)];
      if (L == nz) {                     // top model layer: interface nz+1 and the layer temperature are replaced
        tlev_t = varint(plev_t) + shift;
        tlay = 0.5f * (tlev_t + tlev_b);
      }
    } else {
      plev_t = plev_b - deltap;
      if (L == nlay) plev_t = 0.00f;
      play = 0.5f * (plev_b + plev_t);
      tlev_t = varint(plev_t) + shift;
      tlay = 0.5f * (tlev_t + tlev_b);
    }
    const float pdel = plev_b - plev_t;
    float h2ovmr;
    if (model) { h2ovmr = layer_qv(a.cf, G.at3(i, k, j)) * amdw; h2o_top = h2ovmr; }
    else h2ovmr = h2o_top;
    const float o3mmr = o3_clim(tb, plev_b, plev_t);
    float o3vmr = o3mmr * amdo;
    if (a.o33d && a.o3input == 2) {
      if (model) { o3vmr = a.o33d[G.at3(i, k, j)]; if (L == nz) { o31d_top = o3vmr; o3top_ref = o3mmr; } }
      else {
        o3vmr = o31d_top - o3top_ref * amdo + o3mmr * amdo;
        if (o3vmr <= 0.f) o3vmr = o3mmr * amdo;
      }
    }
    // ---- inatm
    const float amm = (1.f - h2ovmr) * amd + h2ovmr * amw;
    const float coldry = (plev_b - plev_t) * 1.e3f * 6.02214199e+23f / (1.e2f * 9.8066f * amm * (1.f + h2ovmr));
    // summol over imol = 2..7 of the vmr's (co2, o3, n2o, co(=0), ch4, o2) in that order
    float summol = 0.f;
    summol = summol + co2; summol = summol + o3vmr; summol = summol + n2o; summol = summol + 0.f; summol = summol + ch4;
    summol = summol + o2;
    const float wbroad = coldry * (1.f - summol);
    const float wk_h2o = coldry * h2ovmr, wk_co2 = coldry * co2, wk_o3 = coldry * o3vmr, wk_n2o = coldry * n2o,
                wk_co = coldry * 0.f, wk_ch4 = coldry * ch4, wk_o2 = coldry * o2;
    amttl = amttl + coldry + wk_h2o;
    wvttl = wvttl + wk_h2o;
    // ---- setcoef
    PTCoef pc;
    pt_coef(tb.lw_preflog, tb.lw_tref, play, tlay, pc);
    const float water = wk_h2o / coldry;
    const float scalefac = play * stpfac / tlay;
    float forfac, forfrac, selffac, selffrac, scaleminor, scaleminorn2, minorfrac;
    int indfor, indself, indminor;
    if (!(pc.plog <= 4.56f)) {
      laytrop = laytrop + 1;
      forfac = scalefac / (1.f + water);
      float factor = (332.0f - tlay) / 36.0f;
      indfor = min(2, max(1, (int)factor));
      forfrac = factor - (float)indfor;
      selffac = water * forfac;
      factor = (tlay - 188.0f) / 7.2f;
      indself = min(9, max(1, (int)factor - 7));
      selffrac = factor - (float)(indself + 7);
    } else {
      forfac = scalefac / (1.f + water);
      float factor = (tlay - 188.0f) / 36.0f;
      indfor = 3;
      forfrac = factor - 1.0f;
      selffac = water * forfac;
      indself = 0; selffrac = 0.f;
    }
    scaleminor = play / tlay;
    scaleminorn2 = (play / tlay) * (wbroad / (coldry + wk_h2o));
    {
      float factor = (tlay - 180.8f) / 7.2f;
      indminor = min(18, max(1, (int)factor));
      minorfrac = factor - (float)indminor;
    }
    float colh2o = 1.e-20f * wk_h2o, colco2 = 1.e-20f * wk_co2, colo3 = 1.e-20f * wk_o3, coln2o = 1.e-20f * wk_n2o,
          colco = 1.e-20f * wk_co, colch4 = 1.e-20f * wk_ch4, colo2 = 1.e-20f * wk_o2;
    if (colco2 == 0.f) colco2 = 1.e-32f * coldry;
    if (colo3 == 0.f) colo3 = 1.e-32f * coldry;
    if (coln2o == 0.f) coln2o = 1.e-32f * coldry;
    if (colco == 0.f) colco = 1.e-32f * coldry;
    if (colch4 == 0.f) colch4 = 1.e-32f * coldry;
    const float colbrd = 1.e-20f * wbroad;
    selffac = colh2o * selffac;
    forfac = colh2o * forfac;
    COEF(LWC_FAC00, l) = pc.fac00; COEF(LWC_FAC01, l) = pc.fac01; COEF(LWC_FAC10, l) = pc.fac10; COEF(LWC_FAC11, l) = pc.fac11;
    COEF(LWC_H2O, l) = colh2o; COEF(LWC_CO2, l) = colco2; COEF(LWC_O3, l) = colo3; COEF(LWC_N2O, l) = coln2o;
    COEF(LWC_CO, l) = colco; COEF(LWC_CH4, l) = colch4; COEF(LWC_O2, l) = colo2; COEF(LWC_BRD, l) = colbrd;
    COEF(LWC_SELFFAC, l) = selffac; COEF(LWC_SELFFRAC, l) = selffrac; COEF(LWC_FORFAC, l) = forfac; COEF(LWC_FORFRAC, l) = forfrac;
    COEF(LWC_MINORFRAC, l) = minorfrac; COEF(LWC_SCALEMINOR, l) = scaleminor; COEF(LWC_SCALEMINORN2, l) = scaleminorn2;
    COEF(LWC_PAVEL, l) = play; COEF(LWC_COLDRY, l) = coldry; COEF(LWC_TAVEL, l) = tlay; COEF(LWC_TZ, l) = tlev_t;
    COEF(LWC_IDX, l) = __int_as_float(pack_idx(pc.jp, pc.jt, pc.jt1, indself, indfor, indminor));
    if (a.dbg.jp) {
      const size_t q = (size_t)tc * nlay + l;
      a.dbg.jp[q] = pc.jp; a.dbg.jt[q] = pc.jt; a.dbg.jt1[q] = pc.jt1; a.dbg.indfor[q] = indfor; a.dbg.indself[q] = indself;
      a.dbg.indminor[q] = indminor;
      a.dbg.fac00[q] = pc.fac00; a.dbg.fac01[q] = pc.fac01; a.dbg.fac10[q] = pc.fac10; a.dbg.fac11[q] = pc.fac11;
    }
    // ---- aerosol (LW:12576-12615)
    // all 16 band fields of the layer are requested together (one memory latency per layer, not one per band)
    float tv[NBLW];
#pragma unroll
    for (int b = 0; b < NBLW; b++) tv[b] = 0.f;
    if (model && a.aer_ra_feedback == 1) {
      const size_t q = G.at3(i, k, j);
#pragma unroll
      for (int b = 0; b < NBLW; b++) tv[b] = a.tauaerlw[b][q];
      const bool chem = tv[0] > thresh && tv[15] > thresh;
      if (!chem) {
#pragma unroll
        for (int b = 0; b < NBLW; b++) tv[b] = 0.f;
      }
    }
#pragma unroll
    for (int b = 0; b < NBLW; b++) {
      ws.aer[((size_t)b * nlay + l) * cap + c] = tv[b];
      if (model) aersum[b] = aersum[b] + tv[b];
    }
    // ---- clouds
    const bool anycld = (ws.anyc[(size_t)(l >> 5) * cap + c] >> (l & 31)) & 1u;
    if (anycld) {
      LayerCloud lc;
      if (model) layer_cloud(a.cf, G, tb, i, j, k, a.cf.t3d[G.at3(i, k, j)], tlay, pdel, inflg, iceflg, lc);
      else { lc.clwp = 0.f; lc.ciwp = 0.f; lc.cswp = 0.f; lc.rel = 10.f; lc.rei = 10.f; lc.res = 10.f; }
      const float cwp = lc.ciwp + lc.clwp + lc.cswp;
      for (int b = 0; b < NBLW; b++) {
        float tcm = 0.f;
        if (cwp >= 1.e-20f) {
          int rc = lw_cloud_optics(tb, b, iceflg, liqflg, lc.ciwp, lc.clwp, lc.cswp, lc.rei, lc.rel, lc.res, tcm);
          if (rc && !err) err = rc;
        }
        ws.cld[((size_t)b * nlay + l) * cap + c] = tcm;
      }
    }
    plev_b = plev_t; tlev_b = tlev_t;
  }
  if (a.aer_ra_feedback == 1)
    for (int b = 0; b < NBLW; b++) {
      if (aersum[b] < 0.f && !err) err = ARC_ERR_NEG_AOD;
      else if (aersum[b] > 5.f) atomicAdd(a.status + 2, 1);        // "WARNING: Large total lw optical depth" (LW:12616-12627)
    }
  ws.laytrop[c] = laytrop;
  if (a.dbg.laytrop) a.dbg.laytrop[tc] = laytrop;
  // precipitable water and the diffusivity angle per band (inatm LW:11397-11399, rtrnmc LW:3170-3183)
  const float wvsh = (amw * wvttl) / (amd * amttl);
  const float pwvcm = wvsh * (1.e3f * pz0) / (1.e2f * 9.8066f);
  for (int b = 0; b < NBLW; b++) {
    float sd;
    const int ibnd = b + 1;
    if (ibnd == 1 || ibnd == 4 || ibnd >= 10) sd = 1.66f;
    else {
      sd = tb.a0[b] + tb.a1[b] * glm::expf_(tb.a2[b] * pwvcm);
      if (sd > 1.80f) sd = 1.80f;
      if (sd < 1.50f) sd = 1.50f;
    }
    ws.secdiff[(size_t)b * cap + c] = sd;
  }
  set_status(a.status, err);
}
void launch_lw_prep(const LwArgs &a, cudaStream_t s) {
  k_lw_prep<<<(a.ncols + 127) / 128, 128, 0, s>>>(a);
  count_launch();
}

}  // namespace arc
