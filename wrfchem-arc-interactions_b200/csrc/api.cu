// C ABI of the B200 radiation path (include/arc_rad.h): table upload, workspace / staging management, chunked
// kernel pipelines for RRTMG_SWRAD and RRTMG_LWRAD, error reporting.
//
//   arc_rad_init  <- rrtmg_swinit SW:11211 + rrtmg_lwinit LW:12845
//   arc_rad_sw    <- RRTMG_SWRAD  SW:9901        arc_rad_lw <- RRTMG_LWRAD LW:11451
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/arc_rad.h"
#include "glibc_math.cuh"
#include "aer.h"
#include "args.h"

using namespace arc;

namespace {

struct EvPair { std::string name; cudaEvent_t a, b; };

struct Ctx {
  bool ready = false;
  int device = 0;
  cudaStream_t stream = nullptr;
  // second compute stream: the memory-bound sweep + reduce of inner chunk k run here while the compute-bound solver of
  // chunk k+1 runs on `stream` (double-buffered level records); ARC_RAD_OVERLAP=0 puts everything on `stream`
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_solved = nullptr, ev_swept[2] = {nullptr, nullptr};
  bool overlap = true;
  bool bucket = true;           // cloud bucketing of the sunlit column list (ARC_RAD_BUCKET=0: plain tile order)
  // chained LW -> SW step (arc_rad_lwsw with device arrays): 1 = the LW call of the pair (no join, no sync at its end),
  // 2 = the SW call (sunlit compaction, McICA and prep on `stream3` beside the LW kernels; joins and synchronises for both)
  int chain = 0;
  cudaStream_t stream3 = nullptr;
  cudaEvent_t ev_entry = nullptr;      // call entry on `stream`: stream3 starts behind whatever the caller ordered before `stream`
  cudaEvent_t ev_pre = nullptr;
  // asynchronous slab pipeline (host arrays): the chained pair of slab s ends without join / synchronisation, so the LW
  // kernels of slab s+1 start while the last SW sweep of slab s is still running.  ev_lw_done / ev_sw_done (recorded on
  // stream2 after the last LW / SW reduce of a pair) order the reuse of the LW / SW workspaces and the download.
  bool async_pair = false; int slab_index = 0;
  cudaEvent_t ev_lw_done = nullptr, ev_sw_done = nullptr, ev_pre_lw = nullptr;
  HostTables H;
  DevTables D;
  std::vector<void *> table_allocs;
  int *d_status = nullptr, *d_count = nullptr;
  int warn_sw = 0, warn_lw = 0;        // warning counts of the last call (see arc_rad_warning_counts)
  int *d_cols = nullptr; size_t cols_cap = 0;
  int *d_cols_lw = nullptr; size_t cols_lw_cap = 0;      // LW column list (every column, cloud-bucketed)
  // workspaces (grow-only)
  SwWs sw{}; size_t sw_bytes = 0; void *sw_arena = nullptr;
  LwWs lw{}; size_t lw_bytes = 0; void *lw_arena = nullptr;
  // host<->device staging pool (grow-only)
  struct Slot { void *d = nullptr; size_t bytes = 0; };
  std::vector<Slot> pool; size_t pool_next = 0;
  struct Back { void *host; void *dev; size_t bytes; };
  std::vector<Back> backs;
  // pipelined host path: copy streams, events and two staging sets
  cudaStream_t h2d = nullptr, d2h = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  std::vector<Slot> slab[2];
  // timing
  std::vector<EvPair> evs; size_t ev_next = 0;
  std::map<std::string, float> last_ms;
  bool keep_ms = false;        // pipelined path: accumulate kernel times over the slabs of one call
  std::string err;
};
Ctx g;

#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) {                                                                         \
      g.err = std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #call;                   \
      return ARC_ERR_CUDA;                                                                           \
    }                                                                                                \
  } while (0)

template <class T>
int upload(const T *h, size_t n, const T **d) {
  void *p = nullptr;
  CK(cudaMalloc(&p, std::max<size_t>(n * sizeof(T), 16)));
  CK(cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice));
  g.table_allocs.push_back(p);
  *d = (const T *)p;
  return 0;
}
int upload_vec(const std::vector<float> &v, const float **d) { return upload(v.data(), v.size(), d); }

// Two chunk levels: the per-column / per-layer workspace is sized for an outer chunk (ARC_RAD_OUTER, default 262144
// columns: the column-parallel McICA and prep kernels need that many threads to fill the GPU) and the much larger partial
// flux buffer (~0.2 MB per column) for an inner chunk (ARC_RAD_CHUNK, default 32768 columns).
size_t chunk_cap_default() {
  const char *e = getenv("ARC_RAD_CHUNK");
  long v = e ? atol(e) : 32768;
  if (v < 256) v = 256;
  return (size_t)((v + 255) / 256 * 256);
}
size_t outer_cap_default() {
  const char *e = getenv("ARC_RAD_OUTER");
  long v = e ? atol(e) : 262144;
  if (v < 256) v = 256;
  return std::max(chunk_cap_default(), (size_t)((v + 255) / 256 * 256));
}

// LW inner chunk (ARC_RAD_LW_CHUNK, default 32768 columns): bounds the pass records (2 x 141 KB per column at C2's 63 layers)
// and the group-partial buffers (2 x 47 KB per column)
size_t lw_chunk_cap_default() {
  const char *e = getenv("ARC_RAD_LW_CHUNK");
  long v = e ? atol(e) : 32768;
  if (v < 256) v = 256;
  return (size_t)((v + 255) / 256 * 256);
}

// Inner-chunk capacity bounded by the level-record budget (ARC_RAD_REC_GB per spectrum, default 40 GB for the two buffers):
// the records are `bytes_per_col` per column and buffer, e.g. 84 B x 112 g x 52 levels (SW, 3 streams) at C2.
size_t rec_limited_cap(size_t cap, size_t bytes_per_col) {
  const char *e = getenv("ARC_RAD_REC_GB");
  const double budget = (e ? atof(e) : 40.0) * 1e9;
  size_t lim = (size_t)(budget / (2.0 * (double)bytes_per_col));
  lim = std::max<size_t>(256, lim / 256 * 256);
  return std::min(std::min(chunk_cap_default(), cap), lim);
}

struct Carver {
  char *base; size_t off = 0;
  template <class T> T *take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T *p = base ? (T *)(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

void carve_sw(SwWs &w, char *base, size_t &bytes) {
  Carver c{base};
  const size_t cap = w.cap, pcap = w.pcap, nl = w.nlay;
  w.coef = c.take<float>((size_t)SWC_N * nl * cap);
  w.aer = c.take<float>((size_t)NBSW * 3 * nl * cap);
  w.cld = c.take<float>((size_t)NBSW * 4 * nl * cap);
  w.mask = c.take<uint32_t>((size_t)NGSW * w.W * cap);
  w.anyc = c.take<uint32_t>((size_t)w.W * cap);
  w.laytrop = c.take<int>(cap);
  w.laysol = c.take<int>((size_t)NBSW * cap);
  w.colf = c.take<float>((size_t)SWF_N * cap);
  const size_t ns = w.nk / 2;                              // streams = pairs of flux kinds
  w.rec_n = (pcap / REC_TILE) * (nl + 1) * ns * NGSW * SW_REC;   // two buffers of level records (solver k+1 overlaps sweep k)
  w.rec = c.take<float>(2 * w.rec_n);
  w.zinc = c.take<float>((size_t)2 * NGSW * pcap);
  w.bpart = c.take<float>((size_t)sw_sweep_groups() * (nl + 1) * w.nk * pcap);
  w.dirs = c.take<float>((size_t)2 * NGSW * pcap);
  w.uvni = c.take<float>((size_t)sw_sweep_groups() * 2 * pcap);
  bytes = c.off;
}
void carve_lw(LwWs &w, char *base, size_t &bytes) {
  Carver c{base};
  const size_t cap = w.cap, pcap = w.pcap, nl = w.nlay;
  w.coef = c.take<float>((size_t)LWC_N * nl * cap);
  w.aer = c.take<float>((size_t)NBLW * nl * cap);
  w.cld = c.take<float>((size_t)NBLW * nl * cap);
  w.mask = c.take<uint32_t>((size_t)NGLW * w.W * cap);
  w.anyc = c.take<uint32_t>((size_t)w.W * cap);
  w.laytrop = c.take<int>(cap);
  w.colf = c.take<float>((size_t)LWF_N * cap);
  w.secdiff = c.take<float>((size_t)NBLW * cap);
  w.rec = c.take<float4>((size_t)nl * NGLW * pcap);
  w.recC = c.take<float4>((size_t)nl * NGLW * pcap);
  // two buffers of group partials: k_lw_band of inner chunk k+1 runs beside k_lw_reduce of chunk k
  w.bpart = c.take<float>((size_t)2 * lw_sweep_groups() * (nl + 1) * w.nk * pcap);
  bytes = c.off;
}

// slots of the flux kinds in use inside the partial buffer
template <class WS> void set_kinds(WS &w, int variants) {
  int n = 0;
  for (int k = 0; k < NKIND; k++) w.kslot[k] = 0;
  w.kslot[K_FU] = n++; w.kslot[K_FD] = n++; w.kslot[K_CU] = n++; w.kslot[K_CD] = n++;
  if (variants & ARC_VAR_CLEAN) { w.kslot[K_NU] = n++; w.kslot[K_ND] = n++; }
  if (variants & ARC_VAR_CLEANCLEAR) { w.kslot[K_XU] = n++; w.kslot[K_XD] = n++; }
  w.nk = n;
}

// Workspace arena allocation with an error that tells the caller which knobs bound it
int arena_alloc(void **p, size_t need, const char *what) {
  size_t fr = 0, tot = 0;
  cudaMemGetInfo(&fr, &tot);
  cudaError_t e = need <= fr ? cudaMalloc(p, need) : cudaErrorMemoryAllocation;
  if (e != cudaSuccess) {
    cudaGetLastError();
    char msg[320];
    snprintf(msg, sizeof(msg), "%s workspace: %.1f GB needed, %.1f GB free on the device; lower ARC_RAD_CHUNK / ARC_RAD_LW_CHUNK (inner chunk, "
             "columns), ARC_RAD_REC_GB (shortwave level-record budget) or ARC_RAD_OUTER", what, need / 1e9, fr / 1e9);
    g.err = msg;
    return ARC_ERR_CUDA;
  }
  return 0;
}

int ensure_sw_ws(int nlay, size_t cap, size_t pcap, int variants) {
  SwWs w{}; w.cap = (int)cap; w.pcap = (int)pcap; w.nlay = nlay; w.W = (nlay + 31) / 32; set_kinds(w, variants);
  size_t need; carve_sw(w, nullptr, need);
  if (need > g.sw_bytes) {
    if (g.sw_arena) cudaFree(g.sw_arena);
    g.sw_arena = nullptr; g.sw_bytes = 0;
    if (int rc = arena_alloc(&g.sw_arena, need, "shortwave")) return rc;
    g.sw_bytes = need;
  }
  carve_sw(w, (char *)g.sw_arena, need);
  g.sw = w;
  return 0;
}
int ensure_lw_ws(int nlay, size_t cap, size_t pcap, int variants) {
  LwWs w{}; w.cap = (int)cap; w.pcap = (int)pcap; w.nlay = nlay; w.W = (nlay + 31) / 32; set_kinds(w, variants);
  size_t need; carve_lw(w, nullptr, need);
  if (need > g.lw_bytes) {
    if (g.lw_arena) cudaFree(g.lw_arena);
    g.lw_arena = nullptr; g.lw_bytes = 0;
    if (int rc = arena_alloc(&g.lw_arena, need, "longwave")) return rc;
    g.lw_bytes = need;
  }
  carve_lw(w, (char *)g.lw_arena, need);
  g.lw = w;
  return 0;
}

// ---- staging -------------------------------------------------------------------------------------------
int stage_slot(size_t bytes, void **d) {
  if (g.pool_next >= g.pool.size()) g.pool.push_back({});
  Ctx::Slot &s = g.pool[g.pool_next++];
  if (s.bytes < bytes) {
    if (s.d) cudaFree(s.d);
    s.d = nullptr; s.bytes = 0;
    CK(cudaMalloc(&s.d, bytes));
    s.bytes = bytes;
  }
  *d = s.d;
  return 0;
}
// input array: returns the device pointer to use
int in_arr(int memspace, const float *p, size_t n, const float **out) {
  *out = nullptr;
  if (!p) return 0;
  if (memspace == ARC_MEM_DEVICE) { *out = p; return 0; }
  void *d;
  int rc = stage_slot(n * 4, &d);
  if (rc) return rc;
  CK(cudaMemcpyAsync(d, p, n * 4, cudaMemcpyHostToDevice, g.stream));
  *out = (const float *)d;
  return 0;
}
// output array (INOUT semantics: cells outside the tile / night columns keep the caller's values)
int out_arr(int memspace, float *p, size_t n, float **out) {
  *out = nullptr;
  if (!p) return 0;
  if (memspace == ARC_MEM_DEVICE) { *out = p; return 0; }
  void *d;
  int rc = stage_slot(n * 4, &d);
  if (rc) return rc;
  CK(cudaMemcpyAsync(d, p, n * 4, cudaMemcpyHostToDevice, g.stream));
  g.backs.push_back({p, d, n * 4});
  *out = (float *)d;
  return 0;
}
int copy_back() {
  for (auto &b : g.backs) CK(cudaMemcpyAsync(b.host, b.dev, b.bytes, cudaMemcpyDeviceToHost, g.stream));
  g.backs.clear();
  return 0;
}

// ---- timing ----------------------------------------------------------------------------------------------
struct Timed {
  EvPair *e;
  cudaStream_t st;
  explicit Timed(const char *name, cudaStream_t s = nullptr) : st(s ? s : g.stream) {
    if (g.ev_next >= g.evs.size()) {
      EvPair p; p.name = name;
      cudaEventCreate(&p.a); cudaEventCreate(&p.b);
      g.evs.push_back(p);
    }
    e = &g.evs[g.ev_next++];
    e->name = name;
    cudaEventRecord(e->a, st);
  }
  ~Timed() { cudaEventRecord(e->b, st); }
};
void collect_times() {
  for (size_t i = 0; i < g.ev_next; i++) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g.evs[i].a, g.evs[i].b) == cudaSuccess) g.last_ms[g.evs[i].name] += ms;
  }
  g.ev_next = 0;
}

const char *code_msg(int code) {
  switch (code) {
    case ARC_ERR_NEG_AOD: return "ERROR: Negative total optical depth";
    case ARC_ERR_RADIUS: return "ERROR: cloud particle effective size / fdelta out of table bounds";
    case ARC_ERR_UNSUPPORTED: return "unsupported option (iceflag < 3)";
    default: return "device-side error";
  }
}

Geo make_geo(const ArcDims &d) {
  Geo G;
  G.ims = d.ims; G.ime = d.ime; G.kms = d.kms; G.kme = d.kme; G.jms = d.jms; G.jme = d.jme;
  G.its = d.its; G.ite = d.ite; G.jts = d.jts; G.jte = d.jte; G.kts = d.kts; G.kte = d.kte;
  G.ni = d.ime - d.ims + 1; G.nk = d.kme - d.kms + 1;
  G.nci = d.ite - d.its + 1; G.ncol_tile = G.nci * (d.jte - d.jts + 1);
  return G;
}
int check_dims(const ArcDims &d) {
  if (d.its < d.ims || d.ite > d.ime || d.jts < d.jms || d.jte > d.jme || d.kts < d.kms || d.kte + 1 > d.kme ||
      d.ite < d.its || d.jte < d.jts || d.kte - d.kts + 1 < 4) {
    g.err = "bad dimensions: tile must lie inside memory bounds, kte+1 <= kme, at least 4 layers";
    return ARC_ERR_BAD_ARG;
  }
  return 0;
}

template <class T> int dbg_alloc(T *host, size_t n, T **dev, std::vector<std::pair<void *, std::pair<void *, size_t>>> &list) {
  *dev = nullptr;
  if (!host) return 0;
  void *d;
  int rc = stage_slot(n * sizeof(T), &d);
  if (rc) return rc;
  CK(cudaMemsetAsync(d, 0, n * sizeof(T), g.stream));
  list.push_back({host, {d, n * sizeof(T)}});
  *dev = (T *)d;
  return 0;
}

__global__ void k_unpack_mask(const uint32_t *__restrict__ mask, const int *__restrict__ cols, int col0, int ncols, int cap, int W,
                              int nlay, int ngpt, unsigned char *__restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncols) return;
  const int tc = cols ? cols[c] : col0 + c;
  for (int gq = 0; gq < ngpt; gq++)
    for (int l = 0; l < nlay; l++) {
      const uint32_t w = mask[((size_t)gq * W + (l >> 5)) * cap + c];
      out[((size_t)tc * nlay + l) * ngpt + gq] = (w >> (l & 31)) & 1u;
    }
}

int setup_debug(ArcDebug *dbg, size_t ncol, int nlay, int ngpt, DebugTaps &t,
                std::vector<std::pair<void *, std::pair<void *, size_t>>> &list) {
  memset(&t, 0, sizeof(t));
  if (!dbg) return 0;
  int rc = 0;
  const size_t nl = ncol * nlay, ng = nl * ngpt;
  if ((rc = dbg_alloc(dbg->laytrop, ncol, &t.laytrop, list))) return rc;
  // the index taps are written together: require all-or-none
  const bool idx = dbg->jp && dbg->jt && dbg->jt1 && dbg->indfor && dbg->indself && dbg->fac00 && dbg->fac01 && dbg->fac10 && dbg->fac11;
  if (idx) {
    if ((rc = dbg_alloc(dbg->jp, nl, &t.jp, list))) return rc;
    if ((rc = dbg_alloc(dbg->jt, nl, &t.jt, list))) return rc;
    if ((rc = dbg_alloc(dbg->jt1, nl, &t.jt1, list))) return rc;
    if ((rc = dbg_alloc(dbg->indfor, nl, &t.indfor, list))) return rc;
    if ((rc = dbg_alloc(dbg->indself, nl, &t.indself, list))) return rc;
    if ((rc = dbg_alloc(dbg->indminor, nl, &t.indminor, list))) return rc;
    if ((rc = dbg_alloc(dbg->fac00, nl, &t.fac00, list))) return rc;
    if ((rc = dbg_alloc(dbg->fac01, nl, &t.fac01, list))) return rc;
    if ((rc = dbg_alloc(dbg->fac10, nl, &t.fac10, list))) return rc;
    if ((rc = dbg_alloc(dbg->fac11, nl, &t.fac11, list))) return rc;
    if (!t.indminor) {   // kernels write indminor unconditionally with jp: give LW a scratch target
      void *d; if ((rc = stage_slot(nl * 4, &d))) return rc; t.indminor = (int *)d;
    }
  }
  if ((rc = dbg_alloc(dbg->cldmask, ng, &t.cldmask, list))) return rc;
  if (dbg->taug && dbg->taur) {
    if ((rc = dbg_alloc(dbg->taug, ng, &t.taug, list))) return rc;
    if ((rc = dbg_alloc(dbg->taur, ng, &t.taur, list))) return rc;
  }
  if ((rc = dbg_alloc(dbg->sfluxzen, ncol * ngpt, &t.sfluxzen, list))) return rc;
  if ((rc = dbg_alloc(dbg->taucmc, ng, &t.taucmc, list))) return rc;
  if ((rc = dbg_alloc(dbg->hr, nl, &t.hr, list))) return rc;
  return 0;
}

int finish_call(std::vector<std::pair<void *, std::pair<void *, size_t>>> &dbglist) {
  for (auto &e : dbglist) CK(cudaMemcpyAsync(e.first, e.second.first, e.second.second, cudaMemcpyDeviceToHost, g.stream));
  int rc = copy_back();
  if (rc) return rc;
  int st[4] = {0, 0, 0, 0};
  CK(cudaMemcpyAsync(st, g.d_status, sizeof(int) * 4, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaGetLastError());
  collect_times();
  g.warn_sw = st[1]; g.warn_lw = st[2];
  if (st[0]) { g.err = code_msg(st[0]); return st[0]; }
  return 0;
}

void fill_cloud(CloudFields &cf, int memspace, int &rc, int icloud, int warm_rain, int is_cammgmp_used, int has_reqc, int has_reqi,
                int has_reqs, int progn, const int fq[7], float gacc, const float *const p3[18], const float *const p2[3],
                size_t n3, size_t n2) {
  cf.icloud = icloud; cf.warm_rain = warm_rain; cf.is_cammgmp_used = is_cammgmp_used;
  cf.has_reqc = has_reqc; cf.has_reqi = has_reqi; cf.has_reqs = has_reqs; cf.progn = progn;
  cf.f_qv = fq[0]; cf.f_qc = fq[1]; cf.f_qr = fq[2]; cf.f_qi = fq[3]; cf.f_qs = fq[4]; cf.f_qg = fq[5]; cf.f_qndrop = fq[6];
  cf.g = gacc;
  const float **dst3[18] = {&cf.t3d, &cf.cldfra3d, &cf.lradius, &cf.iradius, &cf.qv3d, &cf.qc3d, &cf.qr3d, &cf.qi3d, &cf.qs3d,
                            &cf.qg3d, &cf.qndrop3d, &cf.re_cloud, &cf.re_ice, &cf.re_snow, &cf.f_ice_phy, nullptr, nullptr, nullptr};
  for (int q = 0; q < 15 && !rc; q++) {
    if (q == 9) { cf.qg3d = nullptr; continue; }   // graupel is gathered by the reference but never used (SW:10449)
    rc = in_arr(memspace, p3[q], n3, dst3[q]);
  }
  const float **dst2[3] = {&cf.xland, &cf.xice, &cf.snow};
  for (int q = 0; q < 3 && !rc; q++) rc = in_arr(memspace, p2[q], n2, dst2[q]);
}


// ---------------------------------------------------------------------------------------------------------
// Pipelined host path.  With host arrays the tile is cut into j-slabs (in (i,k,j) order a row range of every field is one
// contiguous block): while slab s computes, slab s+1 is uploaded on a second stream and the outputs of slab s-1 are
// downloaded on a third, through two staging sets.  Each slab runs the ordinary device-memory entry point on staged
// copies whose memory bounds are the slab's rows.  Output arrays are INOUT in the reference (night columns, halo cells and
// levels above kte keep the caller's values): an output is uploaded first only if the call can leave some of its
// staged cells unwritten; otherwise exactly the written sub-block comes back through a strided copy.
enum FieldKind { F3 = 0, F2 = 1, FP = 2 };
struct FieldRef { size_t off; int kind; bool partial; };     // partial: some tile cells may stay unwritten (SW night)
#define FIN(T, f, k) {offsetof(T, f), k, false}
static const FieldRef SW_INS[] = {
    FIN(ArcSwIn, t3d, F3), FIN(ArcSwIn, t8w, F3), FIN(ArcSwIn, p3d, F3), FIN(ArcSwIn, p8w, F3), FIN(ArcSwIn, pi3d, F3),
    FIN(ArcSwIn, cldfra3d, F3), FIN(ArcSwIn, lradius, F3), FIN(ArcSwIn, iradius, F3), FIN(ArcSwIn, qv3d, F3), FIN(ArcSwIn, qc3d, F3),
    FIN(ArcSwIn, qr3d, F3), FIN(ArcSwIn, qi3d, F3), FIN(ArcSwIn, qs3d, F3), FIN(ArcSwIn, qndrop3d, F3), FIN(ArcSwIn, o33d, F3),
    FIN(ArcSwIn, re_cloud, F3), FIN(ArcSwIn, re_ice, F3), FIN(ArcSwIn, re_snow, F3), FIN(ArcSwIn, f_ice_phy, F3),
    FIN(ArcSwIn, tauaer300, F3), FIN(ArcSwIn, tauaer400, F3), FIN(ArcSwIn, tauaer600, F3), FIN(ArcSwIn, tauaer999, F3),
    FIN(ArcSwIn, gaer400, F3), FIN(ArcSwIn, gaer600, F3), FIN(ArcSwIn, waer400, F3), FIN(ArcSwIn, waer600, F3),
    FIN(ArcSwIn, xcoszen, F2), FIN(ArcSwIn, albedo, F2), FIN(ArcSwIn, tsk, F2), FIN(ArcSwIn, xland, F2), FIN(ArcSwIn, xice, F2),
    FIN(ArcSwIn, snow, F2), FIN(ArcSwIn, alswvisdir, F2), FIN(ArcSwIn, alswvisdif, F2), FIN(ArcSwIn, alswnirdir, F2), FIN(ArcSwIn, alswnirdif, F2)};
// inputs the kernels never read: staged as aliases of a neighbour so the "missing field" checks still see them
static const size_t SW_ALIAS[][2] = {{offsetof(ArcSwIn, gaer300), offsetof(ArcSwIn, gaer400)}, {offsetof(ArcSwIn, gaer999), offsetof(ArcSwIn, gaer400)},
                                     {offsetof(ArcSwIn, waer300), offsetof(ArcSwIn, waer400)}, {offsetof(ArcSwIn, waer999), offsetof(ArcSwIn, waer400)}};
#define FOUT(T, f, k, part) {offsetof(T, f), k, part}
static const FieldRef SW_OUTS[] = {
    FOUT(ArcSwOut, rthratensw, F3, true), FOUT(ArcSwOut, gsw, F2, true), FOUT(ArcSwOut, swcf, F2, false), FOUT(ArcSwOut, coszr, F2, false),
    FOUT(ArcSwOut, swupt, F2, false), FOUT(ArcSwOut, swuptc, F2, false), FOUT(ArcSwOut, swuptcln, F2, false), FOUT(ArcSwOut, swdnt, F2, false),
    FOUT(ArcSwOut, swdntc, F2, false), FOUT(ArcSwOut, swdntcln, F2, false), FOUT(ArcSwOut, swupb, F2, false), FOUT(ArcSwOut, swupbc, F2, false),
    FOUT(ArcSwOut, swupbcln, F2, false), FOUT(ArcSwOut, swdnb, F2, false), FOUT(ArcSwOut, swdnbc, F2, false), FOUT(ArcSwOut, swdnbcln, F2, false),
    FOUT(ArcSwOut, swvisdir, F2, false), FOUT(ArcSwOut, swvisdif, F2, false), FOUT(ArcSwOut, swnirdir, F2, false), FOUT(ArcSwOut, swnirdif, F2, false),
    FOUT(ArcSwOut, swddir, F2, false), FOUT(ArcSwOut, swddni, F2, false), FOUT(ArcSwOut, swddif, F2, false),
    FOUT(ArcSwOut, swupflx, FP, true), FOUT(ArcSwOut, swupflxc, FP, true), FOUT(ArcSwOut, swupflxcln, FP, true),
    FOUT(ArcSwOut, swdnflx, FP, true), FOUT(ArcSwOut, swdnflxc, FP, true), FOUT(ArcSwOut, swdnflxcln, FP, true),
    FOUT(ArcSwOut, swuptclnc, F2, false), FOUT(ArcSwOut, swdntclnc, F2, false), FOUT(ArcSwOut, swupbclnc, F2, false), FOUT(ArcSwOut, swdnbclnc, F2, false)};
static const FieldRef LW_INS[] = {
    FIN(ArcLwIn, p8w, F3), FIN(ArcLwIn, p3d, F3), FIN(ArcLwIn, pi3d, F3), FIN(ArcLwIn, t3d, F3), FIN(ArcLwIn, t8w, F3),
    FIN(ArcLwIn, cldfra3d, F3), FIN(ArcLwIn, lradius, F3), FIN(ArcLwIn, iradius, F3), FIN(ArcLwIn, qv3d, F3), FIN(ArcLwIn, qc3d, F3),
    FIN(ArcLwIn, qr3d, F3), FIN(ArcLwIn, qi3d, F3), FIN(ArcLwIn, qs3d, F3), FIN(ArcLwIn, qndrop3d, F3), FIN(ArcLwIn, o33d, F3),
    FIN(ArcLwIn, re_cloud, F3), FIN(ArcLwIn, re_ice, F3), FIN(ArcLwIn, re_snow, F3), FIN(ArcLwIn, f_ice_phy, F3),
    FIN(ArcLwIn, tauaerlw[0], F3), FIN(ArcLwIn, tauaerlw[1], F3), FIN(ArcLwIn, tauaerlw[2], F3), FIN(ArcLwIn, tauaerlw[3], F3),
    FIN(ArcLwIn, tauaerlw[4], F3), FIN(ArcLwIn, tauaerlw[5], F3), FIN(ArcLwIn, tauaerlw[6], F3), FIN(ArcLwIn, tauaerlw[7], F3),
    FIN(ArcLwIn, tauaerlw[8], F3), FIN(ArcLwIn, tauaerlw[9], F3), FIN(ArcLwIn, tauaerlw[10], F3), FIN(ArcLwIn, tauaerlw[11], F3),
    FIN(ArcLwIn, tauaerlw[12], F3), FIN(ArcLwIn, tauaerlw[13], F3), FIN(ArcLwIn, tauaerlw[14], F3), FIN(ArcLwIn, tauaerlw[15], F3),
    FIN(ArcLwIn, emiss, F2), FIN(ArcLwIn, tsk, F2), FIN(ArcLwIn, xland, F2), FIN(ArcLwIn, xice, F2), FIN(ArcLwIn, snow, F2)};
static const FieldRef LW_OUTS[] = {
    FOUT(ArcLwOut, rthratenlw, F3, false), FOUT(ArcLwOut, glw, F2, false), FOUT(ArcLwOut, olr, F2, false), FOUT(ArcLwOut, lwcf, F2, false),
    FOUT(ArcLwOut, lwupt, F2, false), FOUT(ArcLwOut, lwuptc, F2, false), FOUT(ArcLwOut, lwuptcln, F2, false), FOUT(ArcLwOut, lwdnt, F2, false),
    FOUT(ArcLwOut, lwdntc, F2, false), FOUT(ArcLwOut, lwdntcln, F2, false), FOUT(ArcLwOut, lwupb, F2, false), FOUT(ArcLwOut, lwupbc, F2, false),
    FOUT(ArcLwOut, lwupbcln, F2, false), FOUT(ArcLwOut, lwdnb, F2, false), FOUT(ArcLwOut, lwdnbc, F2, false), FOUT(ArcLwOut, lwdnbcln, F2, false),
    FOUT(ArcLwOut, lwupflx, FP, false), FOUT(ArcLwOut, lwupflxc, FP, false), FOUT(ArcLwOut, lwupflxcln, FP, false),
    FOUT(ArcLwOut, lwdnflx, FP, false), FOUT(ArcLwOut, lwdnflxc, FP, false), FOUT(ArcLwOut, lwdnflxcln, FP, false),
    FOUT(ArcLwOut, lwuptclnc, F2, false), FOUT(ArcLwOut, lwdntclnc, F2, false), FOUT(ArcLwOut, lwupbclnc, F2, false), FOUT(ArcLwOut, lwdnbclnc, F2, false)};

template <class T> static inline const float *&fld(T &st, size_t off) { return *reinterpret_cast<const float **>(reinterpret_cast<char *>(&st) + off); }
template <class T> static inline float *&fldw(T &st, size_t off) { return *reinterpret_cast<float **>(reinterpret_cast<char *>(&st) + off); }

static int slab_slot(int set, size_t idx, size_t bytes, void **d) {
  if (g.slab[set].size() <= idx) g.slab[set].resize(idx + 1);
  Ctx::Slot &s = g.slab[set][idx];
  if (s.bytes < bytes) {
    if (s.d) cudaFree(s.d);
    s.d = nullptr; s.bytes = 0;
    CK(cudaMalloc(&s.d, bytes));
    s.bytes = bytes;
  }
  *d = s.d;
  return 0;
}

static int slab_rows_default(int nrows, int ni) {
  const char *e = getenv("ARC_RAD_SLAB_COLUMNS");
  long cols = e ? atol(e) : 32768;
  if (cols <= 0) return nrows;                  // 0: pipelining off
  long r = std::max(1L, cols / std::max(ni, 1));
  return (int)std::min<long>(r, nrows);
}

// One "part" of a pipelined call = one reference entry point (RRTMG_LWRAD or RRTMG_SWRAD) with its argument structs.
// A call may have several parts (arc_rad_lwsw: LW then SW on the same slab); host arrays that several parts share
// (t3d, p3d, qv3d, ... are passed to both adapters by radiation_driver) are uploaded once per slab.
struct PipePart {
  const void *in; size_t in_size;            // host-memspace argument struct (first member: int memspace)
  void *out; size_t out_size;
  const FieldRef *ins; int nins;
  const FieldRef *outs; int nouts;
  const size_t (*alias)[2]; int nalias;
  int (*call)(const ArcDims *, const void *, void *);
};
static inline const float *&fldv(void *st, size_t off) { return *reinterpret_cast<const float **>(reinterpret_cast<char *>(st) + off); }

static int call_sw(const ArcDims *d, const void *in, void *out);
static int call_lw(const ArcDims *d, const void *in, void *out);
static int (*const call_sw_ptr)(const ArcDims *, const void *, void *) = call_sw;
static int (*const call_lw_ptr)(const ArcDims *, const void *, void *) = call_lw;

// Returns -1 when the call is not eligible (too few rows for more than one slab).
static int run_pipelined(const ArcDims &d, const PipePart *parts, int nparts) {
  const int nrows = d.jte - d.jts + 1;
  const int ni = d.ime - d.ims + 1, nk = d.kme - d.kms + 1;
  const int rows_per = slab_rows_default(nrows, d.ite - d.its + 1);
  if (rows_per >= nrows) return -1;
  // Slab boundaries.  The pipeline cannot compute before the first upload has finished and nothing hides the last download, so
  // the slabs ramp up (a quarter, a half, then full slabs of `rows_per` rows: an upload is ~2x faster than the compute of the
  // same rows, so each upload still finishes under the previous slab's compute) and the last 1.5 slabs' worth is cut into a
  // longer and a shorter piece.  Measured on C2 (device timestamps, ARC_RAD_PIPE_TRACE): uniform slabs fill for 5.2 ms and
  // drain for 2.0 ms; ramped ones for 1.5 / 0.9 ms, but six slabs instead of four cost 3.3 ms more compute (every slab pays the
  // latency-bound column kernels and partial waves once): 55.2 -> 54.1 ms per step.  ARC_RAD_SLAB_RAMP=0 restores uniform slabs.
  std::vector<int> slab_j0;      // first row of every slab (relative to jts), plus the end
  {
    const char *e = getenv("ARC_RAD_SLAB_RAMP");
    const bool ramp = !(e && atoi(e) == 0) && nrows >= 3 * rows_per && rows_per >= 8;
    int j = 0;
    if (ramp) {
      for (int r : {rows_per / 4, rows_per / 2}) { slab_j0.push_back(j); j += r; }
      while (nrows - j > rows_per + rows_per / 2) { slab_j0.push_back(j); j += rows_per; }
      const int left = nrows - j, a = std::min(rows_per, (int)(0.65 * left));
      if (a > 0 && left - a > 0) { slab_j0.push_back(j); j += a; }
      slab_j0.push_back(j);
    } else {
      const int n = (nrows + rows_per - 1) / rows_per, each = (nrows + n - 1) / n;
      while (j < nrows) { slab_j0.push_back(j); j = std::min(nrows, j + each); }
    }
    slab_j0.push_back(nrows);
  }
  const int nslab = (int)slab_j0.size() - 1;
  const bool ihalo = d.its != d.ims || d.ite != d.ime;
  const int nz = d.kte - d.kts + 1;
  const size_t row3 = (size_t)ni * nk, row2 = (size_t)ni, rowp = (size_t)ni * (nk + 2);
  auto rowsz = [&](int kind) { return kind == F3 ? row3 : kind == F2 ? row2 : rowp; };
  CK(cudaSetDevice(g.device));
  // staging slot of every field; inputs whose host pointer was already seen (in an earlier part) share that slot
  struct Slotmap { std::vector<int> in_slot, out_slot; };
  std::vector<Slotmap> maps(nparts);
  std::vector<const float *> seen; std::vector<int> seen_slot; std::vector<bool> upload_in;
  int nslots = 0;
  for (int p = 0; p < nparts; p++) {
    maps[p].in_slot.assign(parts[p].nins, -1); maps[p].out_slot.assign(parts[p].nouts, -1);
    for (int f = 0; f < parts[p].nins; f++) {
      const float *h = fldv(const_cast<void *>(parts[p].in), parts[p].ins[f].off);
      if (!h) continue;
      int found = -1;
      for (size_t q = 0; q < seen.size(); q++) if (seen[q] == h) { found = seen_slot[q]; break; }
      if (found < 0) { found = nslots++; seen.push_back(h); seen_slot.push_back(found); }
      maps[p].in_slot[f] = found;
    }
    for (int f = 0; f < parts[p].nouts; f++)
      if (fldv(parts[p].out, parts[p].outs[f].off)) maps[p].out_slot[f] = nslots++;
  }

  auto upload = [&](int s) -> int {
    const int set = s & 1;
    const int j0 = d.jts + slab_j0[s], j1 = d.jts + slab_j0[s + 1] - 1, nr = j1 - j0 + 1;
    CK(cudaStreamWaitEvent(g.h2d, g.ev_out[set], 0));          // the previous user of this set has been downloaded
    std::vector<bool> done(nslots, false);
    for (int p = 0; p < nparts; p++) {
      for (int f = 0; f < parts[p].nins; f++) {
        const int slot = maps[p].in_slot[f];
        if (slot < 0 || done[slot]) continue;
        done[slot] = true;
        const float *h = fldv(const_cast<void *>(parts[p].in), parts[p].ins[f].off);
        const size_t rs = rowsz(parts[p].ins[f].kind);
        void *dv; int rc = slab_slot(set, slot, (size_t)rows_per * rs * 4, &dv); if (rc) return rc;
        CK(cudaMemcpyAsync(dv, h + (size_t)(j0 - d.jms) * rs, (size_t)nr * rs * 4, cudaMemcpyHostToDevice, g.h2d));
      }
      for (int f = 0; f < parts[p].nouts; f++) {
        const int slot = maps[p].out_slot[f];
        if (slot < 0) continue;
        float *h = const_cast<float *>(fldv(parts[p].out, parts[p].outs[f].off));
        const size_t rs = rowsz(parts[p].outs[f].kind);
        void *dv; int rc = slab_slot(set, slot, (size_t)rows_per * rs * 4, &dv); if (rc) return rc;
        if (ihalo || parts[p].outs[f].partial)
          CK(cudaMemcpyAsync(dv, h + (size_t)(j0 - d.jms) * rs, (size_t)nr * rs * 4, cudaMemcpyHostToDevice, g.h2d));
      }
    }
    CK(cudaEventRecord(g.ev_in[set], g.h2d));
    return 0;
  };
  auto download = [&](int s) -> int {
    const int set = s & 1;
    const int j0 = d.jts + slab_j0[s], j1 = d.jts + slab_j0[s + 1] - 1, nr = j1 - j0 + 1;
    for (int p = 0; p < nparts; p++)
      for (int f = 0; f < parts[p].nouts; f++) {
        const int slot = maps[p].out_slot[f];
        if (slot < 0) continue;
        float *h = const_cast<float *>(fldv(parts[p].out, parts[p].outs[f].off));
        const FieldRef &fr = parts[p].outs[f];
        const size_t rs = rowsz(fr.kind);
        const float *dv = (const float *)g.slab[set][slot].d;
        float *hd = h + (size_t)(j0 - d.jms) * rs;
        if (ihalo || fr.partial || fr.kind == F2) {
          CK(cudaMemcpyAsync(hd, dv, (size_t)nr * rs * 4, cudaMemcpyDeviceToHost, g.d2h));
        } else {
          // exactly the written levels of every row: kts..kte (3-D tendency) or kts..kte+2 (flux profile)
          const size_t lev0 = (size_t)(d.kts - d.kms) * ni, nlev = (size_t)(fr.kind == F3 ? nz : nz + 2) * ni;
          CK(cudaMemcpy2DAsync(hd + lev0, rs * 4, dv + lev0, rs * 4, nlev * 4, nr, cudaMemcpyDeviceToHost, g.d2h));
        }
      }
    CK(cudaEventRecord(g.ev_out[set], g.d2h));
    return 0;
  };

  // ARC_RAD_PIPE_TRACE=1: device timestamps of every slab's upload end, last compute kernel and download end (developer aid)
  const bool trace = getenv("ARC_RAD_PIPE_TRACE") != nullptr;
  std::vector<cudaEvent_t> tev;
  auto mark = [&](cudaStream_t st) { if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tev.push_back(e); } };
  mark(g.h2d);
  int rc = upload(0);
  if (rc) return rc;
  mark(g.h2d);
  int status = 0;
  std::vector<std::vector<char>> din(nparts), dout(nparts);
  for (int s = 0; s < nslab; s++) {
    const int set = s & 1;
    if (s + 1 < nslab && (rc = upload(s + 1))) return rc;
    mark(g.h2d);
    const int j0 = d.jts + slab_j0[s], j1 = d.jts + slab_j0[s + 1] - 1;
    ArcDims ds = d;
    ds.jms = j0; ds.jme = j1; ds.jts = j0; ds.jte = j1;
    CK(cudaStreamWaitEvent(g.stream, g.ev_in[set], 0));
    // LW + SW on one slab run as one continuous multi-stream pipeline (see arc_rad_lwsw): no join / sync between the two,
    // and none at the end of the slab either - the next slab's LW kernels start while this slab's last SW sweep runs
    const bool chain = nparts == 2 && g.overlap && parts[0].call == call_lw_ptr && parts[1].call == call_sw_ptr;
    if (chain) CK(cudaStreamWaitEvent(g.stream3, g.ev_in[set], 0));
    for (int p = 0; p < nparts && !status; p++) {
      din[p].assign((const char *)parts[p].in, (const char *)parts[p].in + parts[p].in_size);
      dout[p].assign((const char *)parts[p].out, (const char *)parts[p].out + parts[p].out_size);
      *reinterpret_cast<int *>(din[p].data()) = ARC_MEM_DEVICE;
      for (int f = 0; f < parts[p].nins; f++)
        if (maps[p].in_slot[f] >= 0) fldv(din[p].data(), parts[p].ins[f].off) = (const float *)g.slab[set][maps[p].in_slot[f]].d;
      for (int f = 0; f < parts[p].nouts; f++)
        if (maps[p].out_slot[f] >= 0) fldv(dout[p].data(), parts[p].outs[f].off) = (const float *)g.slab[set][maps[p].out_slot[f]].d;
      for (int q = 0; q < parts[p].nalias; q++)
        if (fldv(din[p].data(), parts[p].alias[q][0])) fldv(din[p].data(), parts[p].alias[q][0]) = fldv(din[p].data(), parts[p].alias[q][1]);
      g.keep_ms = s > 0;
      g.chain = chain ? p + 1 : 0;
      g.async_pair = chain; g.slab_index = s;
      rc = parts[p].call(&ds, din[p].data(), dout[p].data());     // synchronises g.stream before returning (chained: not at all)
      g.chain = 0; g.async_pair = false;
      g.keep_ms = false;
      if (rc && !status) status = rc;
    }
    if (chain && status) { cudaStreamSynchronize(g.stream3); cudaStreamSynchronize(g.stream2); cudaStreamSynchronize(g.stream); collect_times(); }
    if (chain && !status) {       // the download of this slab waits for its last reduce (stream2) and the night-column kernel (stream3)
      CK(cudaStreamWaitEvent(g.d2h, g.ev_sw_done, 0));
      CK(cudaStreamWaitEvent(g.d2h, g.ev_pre, 0));
    }
    if (status) break;            // a failed slab leaves the caller's arrays as they were (nothing of it is copied back)
    mark(g.stream); mark(g.stream2);
    if ((rc = download(s))) return rc;
    mark(g.d2h);
  }
  if (!status && nparts == 2 && g.overlap && parts[0].call == call_lw_ptr && parts[1].call == call_sw_ptr) {
    // end of the asynchronous pipeline: drain the compute streams, fetch the device status word, collect the kernel times
    int st[4] = {0, 0, 0, 0};
    CK(cudaStreamSynchronize(g.stream3)); CK(cudaStreamSynchronize(g.stream2));
    CK(cudaMemcpyAsync(st, g.d_status, sizeof(int) * 4, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    CK(cudaGetLastError());
    collect_times();
    g.warn_sw = st[1]; g.warn_lw = st[2];
    if (st[0]) { g.err = code_msg(st[0]); status = st[0]; }
  }
  CK(cudaStreamSynchronize(g.d2h));
  CK(cudaStreamSynchronize(g.h2d));
  if (trace && tev.size() >= 2) {
    // order of marks: t0, upload(0) end, then per slab: upload(s+1) end, main-stream end, stream2 end, download end
    fprintf(stderr, "pipe trace (ms since the first upload started): upload0 end %.2f", [&] { float m; cudaEventElapsedTime(&m, tev[0], tev[1]); return m; }());
    for (size_t q = 2; q + 3 < tev.size() + 0 && q < tev.size(); q += 4) {
      float m[4] = {0, 0, 0, 0};
      for (int r = 0; r < 4 && q + r < tev.size(); r++) cudaEventElapsedTime(&m[r], tev[0], tev[q + r]);
      fprintf(stderr, " | slab %zu: next upload end %.2f, main stream %.2f, sweep stream %.2f, download end %.2f", (q - 2) / 4, m[0], m[1], m[2], m[3]);
    }
    fprintf(stderr, "\n");
    for (auto e : tev) cudaEventDestroy(e);
  }
  return status;
}

static int call_sw(const ArcDims *d, const void *in, void *out) { return arc_rad_sw_debug(d, (const ArcSwIn *)in, (ArcSwOut *)out, nullptr); }
static int call_lw(const ArcDims *d, const void *in, void *out) { return arc_rad_lw_debug(d, (const ArcLwIn *)in, (ArcLwOut *)out, nullptr); }
static PipePart sw_part(const ArcSwIn *in, ArcSwOut *out) {
  return PipePart{in, sizeof(ArcSwIn), out, sizeof(ArcSwOut), SW_INS, (int)(sizeof(SW_INS) / sizeof(FieldRef)), SW_OUTS,
                  (int)(sizeof(SW_OUTS) / sizeof(FieldRef)), SW_ALIAS, 4, call_sw};
}
static PipePart lw_part(const ArcLwIn *in, ArcLwOut *out) {
  return PipePart{in, sizeof(ArcLwIn), out, sizeof(ArcLwOut), LW_INS, (int)(sizeof(LW_INS) / sizeof(FieldRef)), LW_OUTS,
                  (int)(sizeof(LW_OUTS) / sizeof(FieldRef)), nullptr, 0, call_lw};
}

}  // namespace

// One context (streams, workspaces, status word) per process: the public entry points serialise on a recursive mutex, so calls
// from several host threads (WRF's OpenMP tiles) are safe - they run one after the other (arc_rad_lwsw re-enters arc_rad_lw / _sw).
static std::recursive_mutex g_api_mu;
#define API_LOCK std::lock_guard<std::recursive_mutex> api_lock_(g_api_mu)

extern "C" {

const char *arc_rad_last_error(void) { return g.err.c_str(); }
int arc_rad_lw_nlayers(void) { return g.ready ? g.H.lw_nlayers : 0; }
long long arc_rad_launch_count(void) { return launch_count(); }
void *arc_rad_stream(void) { return (void *)g.stream; }
int arc_rad_test_sweep_groups(const int *ng, int nbands, int gmax, int *band, int *g0, int *size) {
  if (!ng || nbands <= 0 || nbands > 16 || gmax <= 0) return -1;
  int g0s[16], acc = 0, need = 0;
  for (int b = 0; b < nbands; b++) { g0s[b] = acc; acc += ng[b]; need += (ng[b] + gmax - 1) / gmax; }
  if (need > SWEEP_MAXGRP) return -1;
  const SweepGroups G = make_sweep_groups(ng, g0s, nbands, gmax);
  for (int q = 0; q < G.n; q++) { band[q] = G.band[q]; g0[q] = G.g0[q]; size[q] = G.ng[q]; }
  return G.n;
}
long long arc_rad_test_coef_index(int field, int layer, long long column, long long cap, int nfields) {
  return (long long)coef_index(field, layer, (size_t)column, (size_t)cap, nfields);
}
// The reference prints a warning block for every (column, band) whose chem-aerosol column optical depth exceeds 6 in the shortwave
// (where it also rescales the profile to 6, SW:11034-11069) or 5 in the longwave (LW:12616-12627).  Nothing is printed from the
// device: the events of the most recent call (pair) are counted and returned here.
void arc_rad_warning_counts(int *sw_aod_capped, int *lw_aod_large) {
  API_LOCK;
  if (sw_aod_capped) *sw_aod_capped = g.warn_sw;
  if (lw_aod_large) *lw_aod_large = g.warn_lw;
}
int arc_rad_set_overlap(int on) { API_LOCK; const int prev = g.overlap ? 1 : 0; g.overlap = on != 0; return prev; }
float arc_rad_last_kernel_ms(const char *name) {
  auto it = g.last_ms.find(name ? name : "");
  return it == g.last_ms.end() ? -1.f : it->second;
}

void arc_rad_finalize(void) {
  API_LOCK;
  if (!g.ready) return;
  cudaSetDevice(g.device);
  cudaStreamSynchronize(g.stream);
  for (void *p : g.table_allocs) cudaFree(p);
  g.table_allocs.clear();
  if (g.sw_arena) cudaFree(g.sw_arena);
  if (g.lw_arena) cudaFree(g.lw_arena);
  g.sw_arena = g.lw_arena = nullptr; g.sw_bytes = g.lw_bytes = 0;
  for (auto &s : g.pool) if (s.d) cudaFree(s.d);
  g.pool.clear();
  for (int q = 0; q < 2; q++) {
    for (auto &s : g.slab[q]) if (s.d) cudaFree(s.d);
    g.slab[q].clear();
    if (g.ev_in[q]) cudaEventDestroy(g.ev_in[q]);
    if (g.ev_out[q]) cudaEventDestroy(g.ev_out[q]);
    g.ev_in[q] = g.ev_out[q] = nullptr;
  }
  if (g.h2d) cudaStreamDestroy(g.h2d);
  if (g.d2h) cudaStreamDestroy(g.d2h);
  g.h2d = g.d2h = nullptr;
  if (g.d_cols) cudaFree(g.d_cols);
  g.d_cols = nullptr; g.cols_cap = 0;
  if (g.d_cols_lw) cudaFree(g.d_cols_lw);
  g.d_cols_lw = nullptr; g.cols_lw_cap = 0;
  if (g.d_status) cudaFree(g.d_status);
  if (g.d_count) cudaFree(g.d_count);
  g.d_status = g.d_count = nullptr;
  for (auto &e : g.evs) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  g.evs.clear();
  aer_finalize();
  if (g.stream) cudaStreamDestroy(g.stream);
  if (g.stream2) cudaStreamDestroy(g.stream2);
  if (g.stream3) cudaStreamDestroy(g.stream3);
  g.stream = g.stream2 = g.stream3 = nullptr;
  if (g.ev_pre_lw) cudaEventDestroy(g.ev_pre_lw);
  g.ev_pre_lw = nullptr;
  if (g.ev_lw_done) cudaEventDestroy(g.ev_lw_done);
  if (g.ev_sw_done) cudaEventDestroy(g.ev_sw_done);
  g.ev_lw_done = g.ev_sw_done = nullptr;
  if (g.ev_pre) cudaEventDestroy(g.ev_pre);
  g.ev_pre = nullptr;
  if (g.ev_solved) cudaEventDestroy(g.ev_solved);
  for (int q = 0; q < 2; q++) if (g.ev_swept[q]) cudaEventDestroy(g.ev_swept[q]);
  g.ev_solved = g.ev_swept[0] = g.ev_swept[1] = nullptr;
  g.ready = false;
}

int arc_rad_init(const ArcConfig *cfg, const char *sw_data_path, const char *lw_data_path) {
  API_LOCK;
  if (!cfg || !sw_data_path || !lw_data_path) { g.err = "arc_rad_init: null argument"; return ARC_ERR_BAD_ARG; }
  if (g.ready) arc_rad_finalize();
  std::string inl;
  if (cfg->inline_tables) inl = cfg->inline_tables;
  else if (getenv("ARC_RAD_TABLES")) inl = getenv("ARC_RAD_TABLES");
  else { g.err = "arc_rad_init: inline table path missing (ArcConfig.inline_tables or $ARC_RAD_TABLES)"; return ARC_ERR_BAD_ARG; }
  int rc = build_host_tables(inl, sw_data_path, lw_data_path, cfg->cp, cfg->p_top, cfg->kme, g.H, g.err);
  if (rc) return rc;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    g.err = "arc_rad_init: no CUDA device available (this library has no CPU fallback)";
    return ARC_ERR_CUDA;
  }
  if (cfg->device >= 0) { CK(cudaSetDevice(cfg->device)); g.device = cfg->device; }
  else CK(cudaGetDevice(&g.device));
  CK(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
  {
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));     // hi = numerically lowest = greatest priority
    { const char *eb = getenv("ARC_RAD_BUCKET"); g.bucket = !(eb && atoi(eb) == 0); }
    const char *e = getenv("ARC_RAD_OVERLAP");
    g.overlap = !(e && atoi(e) == 0);
    const char *pr = getenv("ARC_RAD_SWEEP_PRIO");
    CK(cudaStreamCreateWithPriority(&g.stream2, cudaStreamNonBlocking, (pr && atoi(pr) == 0) ? lo : hi));
    CK(cudaStreamCreateWithFlags(&g.stream3, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&g.ev_entry, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&g.ev_pre, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&g.ev_lw_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&g.ev_pre_lw, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&g.ev_sw_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&g.ev_solved, cudaEventDisableTiming));
    for (int q = 0; q < 2; q++) CK(cudaEventCreateWithFlags(&g.ev_swept[q], cudaEventDisableTiming));
  }
  CK(cudaStreamCreateWithFlags(&g.h2d, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&g.d2h, cudaStreamNonBlocking));
  for (int q = 0; q < 2; q++) { CK(cudaEventCreateWithFlags(&g.ev_in[q], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&g.ev_out[q], cudaEventDisableTiming)); }
  CK(cudaMalloc(&g.d_status, sizeof(int) * 4));       // [0] first error, [1] SW (column, band) AOD capped at 6, [2] LW (column, band) AOD > 5
  CK(cudaMemset(g.d_status, 0, sizeof(int) * 4));
  CK(cudaMalloc(&g.d_count, 2 * sizeof(int)));

  const HostTables &H = g.H;
  DevTables &D = g.D;
  memset(&D, 0, sizeof(D));
  if ((rc = upload_vec(H.sw_buf, &D.sw_tab))) return rc;
  if ((rc = upload_vec(H.lw_buf, &D.lw_tab))) return rc;
  {
    std::vector<float> e(10004, 0.f);
    for (int i = 0; i < NTBL; i++) e[i] = H.sw_exp_tbl[i];
    if ((rc = upload_vec(e, &D.sw_exp))) return rc;
    std::vector<float> et(2 * 10002, 0.f);
    for (int i = 0; i < NTBL; i++) { et[2 * i] = H.lw_exp_tbl[i]; et[2 * i + 1] = H.lw_tfn_tbl[i]; }
    if ((rc = upload_vec(et, &D.lw_exptfn))) return rc;
  }
  if ((rc = upload_vec(H.get("sw_extliq1"), &D.sw_extliq1))) return rc;
  if ((rc = upload_vec(H.get("sw_ssaliq1"), &D.sw_ssaliq1))) return rc;
  if ((rc = upload_vec(H.get("sw_asyliq1"), &D.sw_asyliq1))) return rc;
  if ((rc = upload_vec(H.get("sw_extice3"), &D.sw_extice3))) return rc;
  if ((rc = upload_vec(H.get("sw_ssaice3"), &D.sw_ssaice3))) return rc;
  if ((rc = upload_vec(H.get("sw_asyice3"), &D.sw_asyice3))) return rc;
  if ((rc = upload_vec(H.get("sw_fdlice3"), &D.sw_fdlice3))) return rc;
  if ((rc = upload_vec(H.get("lw_absliq1"), &D.lw_absliq1))) return rc;
  if ((rc = upload_vec(H.get("lw_absice3"), &D.lw_absice3))) return rc;
  if ((rc = upload_vec(H.get("sw_preflog"), &D.sw_preflog))) return rc;
  if ((rc = upload_vec(H.get("sw_tref"), &D.sw_tref))) return rc;
  if ((rc = upload_vec(H.get("lw_preflog"), &D.lw_preflog))) return rc;
  if ((rc = upload_vec(H.get("lw_tref"), &D.lw_tref))) return rc;
  if ((rc = upload_vec(H.get("lw_chi_mls"), &D.chi_mls))) return rc;
  {
    const std::vector<float> &tp = H.get("lw_totplnk");   // (181,16)
    std::vector<float> pad(184 * 16, 0.f);
    for (int b = 0; b < 16; b++) for (int i = 0; i < 181; i++) pad[184 * b + i] = tp[i + 181 * b];
    if ((rc = upload_vec(pad, &D.totplnk))) return rc;
  }
  {
    // annual-mean ozone profile and half-level pressures, o3data LW:12773-12798 (column independent)
    const std::vector<float> &o3sum = H.get("lw_o3sum"), &ppsum = H.get("lw_ppsum"), &o3win = H.get("lw_o3win"), &ppwin = H.get("lw_ppwin");
    std::vector<float> o3ann(31), ppwrkh(32);
    o3ann[0] = 0.5f * (o3sum[0] + o3win[0]);
    for (int k = 1; k < 31; k++) o3ann[k] = o3win[k - 1] + (o3win[k] - o3win[k - 1]) / (ppwin[k] - ppwin[k - 1]) * (ppsum[k] - ppwin[k - 1]);
    for (int k = 1; k < 31; k++) o3ann[k] = 0.5f * (o3ann[k] + o3sum[k]);
    ppwrkh[0] = 1100.f;
    for (int k = 1; k < 31; k++) ppwrkh[k] = (ppsum[k] + ppsum[k - 1]) / 2.f;
    ppwrkh[31] = 0.f;
    if ((rc = upload_vec(o3ann, &D.o3wrk))) return rc;
    if ((rc = upload_vec(ppwrkh, &D.ppwrkh))) return rc;
  }
  {   // swaerpr (SW:4918-5020): ratio / omega / g of the six ECMWF aerosol types per band, [quantity][type][band]
    std::vector<float> rsr;
    for (const char *n : {"sw_rsrtaua", "sw_rsrpiza", "sw_rsrasya"}) { const std::vector<float> &v = H.get(n); rsr.insert(rsr.end(), v.begin(), v.end()); }
    if (rsr.size() != 3 * 6 * 14) { g.err = "inline tables: swaerpr arrays missing"; return ARC_ERR_IO; }
    if ((rc = upload_vec(rsr, &D.sw_rsr))) return rc;
  }
  if ((rc = upload_vec(H.get("lw_retab"), &D.retab))) return rc;
  if ((rc = upload_vec(H.get("lw_pprof"), &D.pprof))) return rc;
  if ((rc = upload_vec(H.get("lw_tprof"), &D.tprof))) return rc;
  D.heatfac = H.heatfac; D.fluxfac = H.fluxfac; D.oneminus = H.oneminus; D.bpade = H.bpade;
  {
    const std::vector<float> &wmin = H.get("sw_wavemin"), &wmax = H.get("sw_wavemax");
    for (int b = 0; b < 14; b++) D.wavemid[b] = 0.5f * (wmin[b] + wmax[b]);
    const std::vector<float> &a0 = H.get("lw_a0"), &a1 = H.get("lw_a1"), &a2 = H.get("lw_a2");
    for (int b = 0; b < 16; b++) { D.a0[b] = a0[b]; D.a1[b] = a1[b]; D.a2[b] = a2[b]; D.delwave[b] = H.lw_delwave[b]; }
  }
  D.lw_nlayers = H.lw_nlayers;
  upload_band_descs(H);
  if (!lw_layout_ok()) { g.err = "RRTMG_LW_DATA: table shapes differ from the RRTMG layout the longwave kernel is compiled for"; return ARC_ERR_IO; }
  CK(cudaDeviceSynchronize());
  g.ready = true;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
int arc_rad_sw_debug(const ArcDims *d, const ArcSwIn *in, ArcSwOut *out, ArcDebug *dbg) {
  API_LOCK;
  if (!g.ready) { g.err = "arc_rad_sw: not initialised"; return ARC_ERR_NOT_INIT; }
  if (!d || !in || !out) { g.err = "arc_rad_sw: null argument"; return ARC_ERR_BAD_ARG; }
  int rc = check_dims(*d);
  if (rc) return rc;
  if (in->memspace == ARC_MEM_HOST && !dbg && !in->tauaer3d_sw && in->aer_opt != 1) {
    const PipePart part = sw_part(in, out);
    rc = run_pipelined(*d, &part, 1);
    if (rc != -1) return rc;
    rc = 0;
  }
  // argument checks of the reference (SW:10288-10305, chemics_init.F:406-408)
  if (in->aer_ra_feedback == 1 &&
      !(in->tauaer300 && in->tauaer400 && in->tauaer600 && in->tauaer999 && in->gaer300 && in->gaer400 && in->gaer600 &&
        in->gaer999 && in->waer300 && in->waer400 && in->waer600 && in->waer999)) {
    g.err = "Warning: missing fields required for aerosol radiation"; return ARC_ERR_MISSING_FIELD;
  }
  if (in->clean_atm_diag > 0 && in->aer_ra_feedback <= 0) {
    g.err = "clean_atm_diag > 0 requires aer_ra_feedback > 0 (chemics_init.F:406-408)"; return ARC_ERR_CONFIG;
  }
  if (in->aer_opt == 1) {
    // iaer = 6 (SW:9201-9205): the six ECMWF aerosol types with the optical depths AEROD(i,k,j,1:6).  The reference's second
    // ("clean") spcvmc_sw call reads ztauacln, which only the iaer = 10 branch defines (SW:9343-9352): undefined there, refused here.
    if (!in->aerod || in->no_src < 6) { g.err = "aer_opt=1 needs aerod(i,k,j,1:6) (no_src >= 6)"; return ARC_ERR_MISSING_FIELD; }
    if (in->clean_atm_diag > 0) {
      g.err = "aer_opt=1 with clean_atm_diag: the reference leaves the clean call's aerosol optical depth undefined"; return ARC_ERR_UNSUPPORTED;
    }
  }
  if (!in->xcoszen || !in->albedo || !in->t3d || !in->t8w || !in->p3d || !in->p8w || !in->pi3d || !in->qv3d || !in->xland ||
      !in->xice || !in->snow || !out->rthratensw || !out->gsw || !out->swcf || !out->coszr || !out->swddir || !out->swddni ||
      !out->swddif) {
    g.err = "arc_rad_sw: required array missing"; return ARC_ERR_BAD_ARG;
  }
  if (in->icloud != 0 && ((in->has_reqc && !in->re_cloud) || (in->has_reqi && !in->re_ice) || (in->has_reqs && !in->re_snow))) {
    g.err = "arc_rad_sw: has_req* set but re_* array missing"; return ARC_ERR_BAD_ARG;
  }
  if (in->sf_surface_physics == 8 && !(in->alswvisdir && in->alswvisdif && in->alswnirdir && in->alswnirdif)) {
    g.err = "arc_rad_sw: SSiB albedos missing"; return ARC_ERR_BAD_ARG;
  }
  if ((out->swupflx || out->swupflxc || out->swupflxcln || out->swdnflx || out->swdnflxc || out->swdnflxcln) &&
      !(out->swupflx && out->swupflxc && out->swupflxcln && out->swdnflx && out->swdnflxc && out->swdnflxcln)) {
    g.err = "arc_rad_sw: flux profile outputs must be passed all together"; return ARC_ERR_BAD_ARG;
  }
  {   // optional output groups are all-or-nothing: every array the caller passes is written (the host path relies on it)
    float *const grp[16] = {out->swupt, out->swuptc, out->swuptcln, out->swdnt, out->swdntc, out->swdntcln, out->swupb, out->swupbc,
                            out->swupbcln, out->swdnb, out->swdnbc, out->swdnbcln, out->swvisdir, out->swvisdif, out->swnirdir, out->swnirdif};
    int n = 0; for (float *p : grp) n += p != nullptr;
    if (n != 0 && n != 16) { g.err = "arc_rad_sw: TOA/surface flux outputs must be passed all together"; return ARC_ERR_BAD_ARG; }
    float *const ex[4] = {out->swuptclnc, out->swdntclnc, out->swupbclnc, out->swdnbclnc};
    n = 0; for (float *p : ex) n += p != nullptr;
    if (n != 0 && n != 4) { g.err = "arc_rad_sw: the four clean-clear outputs must be passed all together"; return ARC_ERR_BAD_ARG; }
  }
  CK(cudaSetDevice(g.device));
  if (!g.keep_ms) for (auto it = g.last_ms.begin(); it != g.last_ms.end();) { if (it->first.compare(0, 3, "sw_") == 0) it = g.last_ms.erase(it); else ++it; }
  g.pool_next = 0; g.backs.clear();
  const int ms = in->memspace;
  SwArgs a{};
  a.geo = make_geo(*d);
  a.tb = g.D;
  const Geo &G = a.geo;
  const size_t n3 = G.n3(), n2 = G.n2(), np = G.np();
  const int nz = d->kte - d->kts + 1, nlay = nz + 1;
  if (nlay > 159) { g.err = "arc_rad_sw: too many layers (max 158 model layers)"; return ARC_ERR_BAD_ARG; }

  const bool chained = g.chain == 2;
  cudaStream_t sp = chained ? g.stream3 : g.stream;       // stream of the column-parallel pre-kernels of the first outer chunk
  if (!chained) CK(cudaMemsetAsync(g.d_status, 0, sizeof(int) * 4, g.stream));
  if (chained && g.async_pair) CK(cudaStreamWaitEvent(g.stream3, g.ev_sw_done, 0));    // the SW workspace of the previous slab is free
  {
    const int fq[7] = {in->f_qv, in->f_qc, in->f_qr, in->f_qi, in->f_qs, in->f_qg, in->f_qndrop};
    const float *const p3[18] = {in->t3d, in->cldfra3d, in->lradius, in->iradius, in->qv3d, in->qc3d, in->qr3d, in->qi3d, in->qs3d,
                                 in->qg3d, in->qndrop3d, in->re_cloud, in->re_ice, in->re_snow, in->f_ice_phy, 0, 0, 0};
    const float *const p2[3] = {in->xland, in->xice, in->snow};
    fill_cloud(a.cf, ms, rc, in->icloud, in->warm_rain, in->is_cammgmp_used, in->has_reqc, in->has_reqi, in->has_reqs, in->progn, fq,
               in->g, p3, p2, n3, n2);
    if (rc) return rc;
  }
  a.o3input = in->o3input; a.aer_ra_feedback = in->aer_ra_feedback; a.sf_surface_physics = in->sf_surface_physics;
  a.solcon = in->solcon;
#define IN3(f) if ((rc = in_arr(ms, in->f, n3, &a.f))) return rc
#define IN2(f) if ((rc = in_arr(ms, in->f, n2, &a.f))) return rc
  IN3(t8w); IN3(p3d); IN3(p8w); IN3(pi3d); IN3(o33d);
  IN2(tsk);
  if (in->aer_ra_feedback == 1) { IN3(tauaer300); IN3(tauaer400); IN3(tauaer600); IN3(tauaer999); IN3(gaer400); IN3(gaer600); IN3(waer400); IN3(waer600); }
  if (in->aer_opt == 1 && (rc = in_arr(ms, in->aerod, n3 * 6, &a.aerod))) return rc;
  if (in->tauaer3d_sw && in->ssaaer3d_sw && in->asyaer3d_sw) {
    if ((rc = in_arr(ms, in->tauaer3d_sw, n3 * 14, &a.tauaer3d_sw))) return rc;
    if ((rc = in_arr(ms, in->ssaaer3d_sw, n3 * 14, &a.ssaaer3d_sw))) return rc;
    if ((rc = in_arr(ms, in->asyaer3d_sw, n3 * 14, &a.asyaer3d_sw))) return rc;
  }
  IN2(xcoszen); IN2(albedo);
  if (in->sf_surface_physics == 8) { IN2(alswvisdir); IN2(alswvisdif); IN2(alswnirdir); IN2(alswnirdif); }
#undef IN3
#undef IN2
#define OUT3(f) if ((rc = out_arr(ms, out->f, n3, &a.f))) return rc
#define OUT2(f) if ((rc = out_arr(ms, out->f, n2, &a.f))) return rc
#define OUTP(f) if ((rc = out_arr(ms, out->f, np, &a.f))) return rc
  OUT3(rthratensw); OUT2(gsw); OUT2(swcf); OUT2(coszr);
  OUT2(swupt); OUT2(swuptc); OUT2(swuptcln); OUT2(swdnt); OUT2(swdntc); OUT2(swdntcln);
  OUT2(swupb); OUT2(swupbc); OUT2(swupbcln); OUT2(swdnb); OUT2(swdnbc); OUT2(swdnbcln);
  OUT2(swvisdir); OUT2(swvisdif); OUT2(swnirdir); OUT2(swnirdif); OUT2(swddir); OUT2(swddni); OUT2(swddif);
  OUTP(swupflx); OUTP(swupflxc); OUTP(swupflxcln); OUTP(swdnflx); OUTP(swdnflxc); OUTP(swdnflxcln);
  const bool ext = out->swuptclnc && out->swdntclnc && out->swupbclnc && out->swdnbclnc;
  if (ext) { OUT2(swuptclnc); OUT2(swdntclnc); OUT2(swupbclnc); OUT2(swdnbclnc); }
#undef OUT3
#undef OUT2
#undef OUTP
  int variants = in->variant_mask;
  if (variants == 0) variants = ARC_VAR_FULL | ARC_VAR_CLEAR | (in->clean_atm_diag > 0 ? ARC_VAR_CLEAN : 0);
  variants |= ARC_VAR_FULL | ARC_VAR_CLEAR;
  if (in->clean_atm_diag <= 0) variants &= ~(ARC_VAR_CLEAN | ARC_VAR_CLEANCLEAR);
  if (ext && (variants & ARC_VAR_CLEAN)) variants |= ARC_VAR_CLEANCLEAR;
  if (!ext) variants &= ~ARC_VAR_CLEANCLEAR;
  a.variants = variants;
  a.ngroups = sw_sweep_groups();
  a.status = g.d_status;

  std::vector<std::pair<void *, std::pair<void *, size_t>>> dbglist;
  if ((rc = setup_debug(dbg, (size_t)G.ncol_tile, nlay, NGSW, a.dbg, dbglist))) return rc;

  // sunlit compaction
  if ((size_t)G.ncol_tile > g.cols_cap) {
    if (g.d_cols) cudaFree(g.d_cols);
    g.d_cols = nullptr; g.cols_cap = 0;
    CK(cudaMalloc(&g.d_cols, sizeof(int) * (size_t)G.ncol_tile));
    g.cols_cap = (size_t)G.ncol_tile;
  }
  int nsun = 0;
  {
    Timed t("sw_compact", sp);
    launch_compact_sunlit(G, a.xcoszen, (in->icloud != 0 && g.bucket) ? a.cf.cldfra3d : nullptr, g.d_cols, g.d_count, 0, sp);
    launch_sw_night(a, sp);
  }
  CK(cudaMemcpyAsync(&nsun, g.d_count, sizeof(int), cudaMemcpyDeviceToHost, sp));
  CK(cudaStreamSynchronize(sp));
  if (nsun > 0) {
    const size_t cap = std::min(outer_cap_default(), (size_t)((nsun + 255) / 256 * 256));
    const int nstream = 2 + ((variants & ARC_VAR_CLEAN) ? 1 : 0) + ((variants & ARC_VAR_CLEANCLEAR) ? 1 : 0);
    const size_t pcap = rec_limited_cap(cap, (size_t)28 * nstream * NGSW * (nlay + 1));
    if ((rc = ensure_sw_ws(nlay, cap, pcap, variants))) return rc;
    for (int o0 = 0; o0 < nsun; o0 += (int)cap) {
      const int no = std::min((int)cap, nsun - o0);
      a.ws = g.sw;
      a.ws.cols = g.d_cols + o0;
      a.ncols = no;
      McicaArgs m{};
      m.geo = G; m.nlay = nlay; m.nz = nz; m.ngpt = NGSW; m.permuteseed = 1; m.ncols = no; m.W = a.ws.W; m.icloud = in->icloud;
      m.cap = (int)cap; m.col0 = 0; m.lw_buffer = 0; m.cols = a.ws.cols; m.p3d = a.p3d; m.p8w = a.p8w; m.cldfra3d = a.cf.cldfra3d;
      m.mask = a.ws.mask; m.anyc = a.ws.anyc;
      cudaStream_t so = o0 == 0 ? sp : g.stream;
      { Timed t("sw_mcica", so); launch_mcica(m, so); }
      { Timed t("sw_prep", so); launch_sw_prep(a, so); }
      if (a.dbg.cldmask) {
        k_unpack_mask<<<(no + 127) / 128, 128, 0, so>>>(a.ws.mask, a.ws.cols, 0, no, (int)cap, a.ws.W, nlay, NGSW, a.dbg.cldmask);
        count_launch();
      }
      if (so != g.stream) { CK(cudaEventRecord(g.ev_pre, so)); CK(cudaStreamWaitEvent(g.stream, g.ev_pre, 0)); }
      // inner chunks: column-indexed workspace pointers advance by c0, the partial buffers restart at 0
      cudaStream_t s2 = g.overlap ? g.stream2 : g.stream;
      int kc = 0;
      for (int c0 = 0; c0 < no; c0 += (int)pcap, kc++) {
        SwArgs b = a;
        b.ncols = std::min((int)pcap, no - c0);
        b.ws.cols += c0; b.ws.coef += (size_t)c0 * SWC_N; b.ws.aer += c0; b.ws.cld += c0; b.ws.mask += c0; b.ws.anyc += c0;
        b.ws.laytrop += c0; b.ws.laysol += c0; b.ws.colf += c0;
        const int buf = kc & 1;
        b.ws.rec += buf * a.ws.rec_n;
        b.ws.zinc += (size_t)buf * NGSW * pcap; b.ws.dirs += (size_t)buf * NGSW * pcap;
        if (g.overlap && kc >= 2) CK(cudaStreamWaitEvent(g.stream, g.ev_swept[buf], 0));      // records of chunk k-2 consumed
        { Timed t("sw_solve"); launch_sw_solve(b, g.stream); }
        if (g.overlap) { CK(cudaEventRecord(g.ev_solved, g.stream)); CK(cudaStreamWaitEvent(s2, g.ev_solved, 0)); }
        { Timed t("sw_sweep", s2); launch_sw_sweep(b, s2); }
        { Timed t("sw_reduce", s2); launch_sw_reduce(b, s2); }
        if (g.overlap) CK(cudaEventRecord(g.ev_swept[buf], s2));
      }
      if (g.overlap) { CK(cudaStreamWaitEvent(g.stream, g.ev_swept[0], 0)); CK(cudaStreamWaitEvent(g.stream, g.ev_swept[1], 0)); }
    }
  }
  if (chained && g.async_pair) {   // slab pipeline: no join here; run_pipelined orders the download on these events
    CK(cudaEventRecord(g.ev_pre, g.stream3));
    CK(cudaEventRecord(g.ev_sw_done, g.stream2));
    return 0;
  }
  if (chained) {   // join everything the pair queued: the night-column kernel on stream3, the sweeps of both calls on stream2
    CK(cudaEventRecord(g.ev_pre, g.stream3)); CK(cudaStreamWaitEvent(g.stream, g.ev_pre, 0));
    CK(cudaStreamWaitEvent(g.stream, g.ev_swept[0], 0)); CK(cudaStreamWaitEvent(g.stream, g.ev_swept[1], 0));
  }
  return finish_call(dbglist);
}

int arc_rad_sw(const ArcDims *d, const ArcSwIn *in, ArcSwOut *out) { return arc_rad_sw_debug(d, in, out, nullptr); }

// ---------------------------------------------------------------------------------------------------------
int arc_rad_lw_debug(const ArcDims *d, const ArcLwIn *in, ArcLwOut *out, ArcDebug *dbg) {
  API_LOCK;
  if (!g.ready) { g.err = "arc_rad_lw: not initialised"; return ARC_ERR_NOT_INIT; }
  if (!d || !in || !out) { g.err = "arc_rad_lw: null argument"; return ARC_ERR_BAD_ARG; }
  int rc = check_dims(*d);
  if (rc) return rc;
  if (in->memspace == ARC_MEM_HOST && !dbg) {
    const PipePart part = lw_part(in, out);
    rc = run_pipelined(*d, &part, 1);
    if (rc != -1) return rc;
    rc = 0;
  }
  if (in->aer_ra_feedback == 1)
    for (int b = 0; b < 16; b++)
      if (!in->tauaerlw[b]) { g.err = "Warning: missing fields required for aerosol radiation"; return ARC_ERR_MISSING_FIELD; }
  if (in->clean_atm_diag > 0 && in->aer_ra_feedback <= 0) {
    g.err = "clean_atm_diag > 0 requires aer_ra_feedback > 0 (chemics_init.F:406-408)"; return ARC_ERR_CONFIG;
  }
  if (!in->emiss || !in->t3d || !in->t8w || !in->p3d || !in->p8w || !in->pi3d || !in->qv3d || !in->tsk || !in->xland || !in->xice ||
      !in->snow || !out->rthratenlw || !out->glw || !out->olr || !out->lwcf) {
    g.err = "arc_rad_lw: required array missing"; return ARC_ERR_BAD_ARG;
  }
  if (in->icloud != 0 && ((in->has_reqc && !in->re_cloud) || (in->has_reqi && !in->re_ice) || (in->has_reqs && !in->re_snow))) {
    g.err = "arc_rad_lw: has_req* set but re_* array missing"; return ARC_ERR_BAD_ARG;
  }
  if ((out->lwupflx || out->lwupflxc || out->lwupflxcln || out->lwdnflx || out->lwdnflxc || out->lwdnflxcln) &&
      !(out->lwupflx && out->lwupflxc && out->lwupflxcln && out->lwdnflx && out->lwdnflxc && out->lwdnflxcln)) {
    g.err = "arc_rad_lw: flux profile outputs must be passed all together"; return ARC_ERR_BAD_ARG;
  }
  {   // optional output groups are all-or-nothing (see arc_rad_sw)
    float *const grp[12] = {out->lwupt, out->lwuptc, out->lwuptcln, out->lwdnt, out->lwdntc, out->lwdntcln, out->lwupb, out->lwupbc,
                            out->lwupbcln, out->lwdnb, out->lwdnbc, out->lwdnbcln};
    int n = 0; for (float *p : grp) n += p != nullptr;
    if (n != 0 && n != 12) { g.err = "arc_rad_lw: TOA/surface flux outputs must be passed all together"; return ARC_ERR_BAD_ARG; }
    float *const ex[4] = {out->lwuptclnc, out->lwdntclnc, out->lwupbclnc, out->lwdnbclnc};
    n = 0; for (float *p : ex) n += p != nullptr;
    if (n != 0 && n != 4) { g.err = "arc_rad_lw: the four clean-clear outputs must be passed all together"; return ARC_ERR_BAD_ARG; }
  }
  if (in->variant_mask != 0 && (in->variant_mask & ~(ARC_VAR_FULL | ARC_VAR_CLEAR | ARC_VAR_CLEAN | ARC_VAR_CLEANCLEAR))) {
    g.err = "arc_rad_lw: unknown bits in variant_mask"; return ARC_ERR_BAD_ARG;
  }
  CK(cudaSetDevice(g.device));
  if (!g.keep_ms) for (auto it = g.last_ms.begin(); it != g.last_ms.end();) { if (it->first.compare(0, 3, "lw_") == 0) it = g.last_ms.erase(it); else ++it; }
  g.pool_next = 0; g.backs.clear();
  const int ms = in->memspace;
  LwArgs a{};
  a.geo = make_geo(*d);
  a.tb = g.D;
  const Geo &G = a.geo;
  const size_t n3 = G.n3(), n2 = G.n2(), np = G.np();
  const int nz = d->kte - d->kts + 1;
  // nlay = nlayers = kme + nint(p_top/4 hPa) - 1 as rrtmg_lwinit fixed it (LW:12861, 12050): the reference's layer count does
  // not depend on the tile, which presumes the WRF convention kts = 1
  if (d->kts != 1) { g.err = "arc_rad_lw: kts must be 1 (the layer count nlayers of rrtmg_lwinit assumes it, LW:12861)"; return ARC_ERR_BAD_ARG; }
  const int nlay = g.H.lw_nlayers;
  if (nlay > 159 || nlay < nz + 1) { g.err = "arc_rad_lw: bad LW layer count (nlayers from init vs kte)"; return ARC_ERR_BAD_ARG; }

  if (!(g.chain == 1 && g.async_pair && g.slab_index > 0)) CK(cudaMemsetAsync(g.d_status, 0, sizeof(int) * 4, g.stream));
  // Slab pipeline: from the second slab on the LW column kernels (McICA, prep) run on stream3 beside the SW solver of the
  // previous slab (at one slab's worth of columns they are latency-bound and would otherwise sit exposed on the main stream)
  cudaStream_t spl = (g.chain == 1 && g.async_pair && g.slab_index > 0) ? g.stream3 : g.stream;
  if (g.chain == 1 && g.async_pair) CK(cudaStreamWaitEvent(spl, g.ev_lw_done, 0));     // the LW workspace of the previous slab is free
  {
    const int fq[7] = {in->f_qv, in->f_qc, in->f_qr, in->f_qi, in->f_qs, in->f_qg, in->f_qndrop};
    const float *const p3[18] = {in->t3d, in->cldfra3d, in->lradius, in->iradius, in->qv3d, in->qc3d, in->qr3d, in->qi3d, in->qs3d,
                                 in->qg3d, in->qndrop3d, in->re_cloud, in->re_ice, in->re_snow, in->f_ice_phy, 0, 0, 0};
    const float *const p2[3] = {in->xland, in->xice, in->snow};
    fill_cloud(a.cf, ms, rc, in->icloud, in->warm_rain, in->is_cammgmp_used, in->has_reqc, in->has_reqi, in->has_reqs, in->progn, fq,
               in->g, p3, p2, n3, n2);
    if (rc) return rc;
  }
  a.o3input = in->o3input; a.aer_ra_feedback = in->aer_ra_feedback;
#define IN3(f) if ((rc = in_arr(ms, in->f, n3, &a.f))) return rc
#define IN2(f) if ((rc = in_arr(ms, in->f, n2, &a.f))) return rc
  IN3(t8w); IN3(p3d); IN3(p8w); IN3(pi3d); IN3(o33d);
  IN2(tsk); IN2(emiss);
  if (in->aer_ra_feedback == 1)
    for (int b = 0; b < 16; b++) if ((rc = in_arr(ms, in->tauaerlw[b], n3, &a.tauaerlw[b]))) return rc;
#undef IN3
#undef IN2
#define OUT3(f) if ((rc = out_arr(ms, out->f, n3, &a.f))) return rc
#define OUT2(f) if ((rc = out_arr(ms, out->f, n2, &a.f))) return rc
#define OUTP(f) if ((rc = out_arr(ms, out->f, np, &a.f))) return rc
  OUT3(rthratenlw); OUT2(glw); OUT2(olr); OUT2(lwcf);
  OUT2(lwupt); OUT2(lwuptc); OUT2(lwuptcln); OUT2(lwdnt); OUT2(lwdntc); OUT2(lwdntcln);
  OUT2(lwupb); OUT2(lwupbc); OUT2(lwupbcln); OUT2(lwdnb); OUT2(lwdnbc); OUT2(lwdnbcln);
  OUTP(lwupflx); OUTP(lwupflxc); OUTP(lwupflxcln); OUTP(lwdnflx); OUTP(lwdnflxc); OUTP(lwdnflxcln);
  const bool ext = out->lwuptclnc && out->lwdntclnc && out->lwupbclnc && out->lwdnbclnc;
  if (ext) { OUT2(lwuptclnc); OUT2(lwdntclnc); OUT2(lwupbclnc); OUT2(lwdnbclnc); }
#undef OUT3
#undef OUT2
#undef OUTP
  // call variants: as arc_rad_sw (full + clear always; clean from clean_atm_diag, LW:11022-11027, unless variant_mask narrows it)
  int variants = in->variant_mask;
  if (variants == 0) variants = ARC_VAR_FULL | ARC_VAR_CLEAR | (in->clean_atm_diag > 0 ? ARC_VAR_CLEAN : 0);
  variants |= ARC_VAR_FULL | ARC_VAR_CLEAR;
  if (in->clean_atm_diag <= 0) variants &= ~(ARC_VAR_CLEAN | ARC_VAR_CLEANCLEAR);
  if (ext && (variants & ARC_VAR_CLEAN)) variants |= ARC_VAR_CLEANCLEAR;
  if (!ext) variants &= ~ARC_VAR_CLEANCLEAR;
  a.variants = variants;
  a.ngroups = lw_sweep_groups();
  a.status = g.d_status;

  std::vector<std::pair<void *, std::pair<void *, size_t>>> dbglist;
  if ((rc = setup_debug(dbg, (size_t)G.ncol_tile, nlay, NGLW, a.dbg, dbglist))) return rc;

  const int ncol = G.ncol_tile;
  const size_t cap = std::min(outer_cap_default(), (size_t)((ncol + 255) / 256 * 256));
  const size_t pcap = std::min(lw_chunk_cap_default(), cap);
  if ((rc = ensure_lw_ws(nlay, cap, pcap, variants))) return rc;
  // column order: cloud-free columns first, then the columns with cloud (see launch_compact_sunlit): the cloudy-layer branch of
  // rtrnmc then runs in warps whose columns all have cloud instead of in every warp that holds one such column
  const bool lw_bucket = g.bucket && in->icloud != 0 && a.cf.cldfra3d != nullptr;
  if (lw_bucket) {
    if ((size_t)ncol > g.cols_lw_cap) {
      if (g.d_cols_lw) cudaFree(g.d_cols_lw);
      g.d_cols_lw = nullptr; g.cols_lw_cap = 0;
      CK(cudaMalloc(&g.d_cols_lw, sizeof(int) * (size_t)ncol));
      g.cols_lw_cap = (size_t)ncol;
    }
    launch_compact_sunlit(G, nullptr, a.cf.cldfra3d, g.d_cols_lw, g.d_count + 1, 1, spl);
  }
  for (int o0 = 0; o0 < ncol; o0 += (int)cap) {
    const int no = std::min((int)cap, ncol - o0);
    a.ws = g.lw;
    a.ws.cols = lw_bucket ? g.d_cols_lw + o0 : nullptr;
    a.col0 = o0;
    a.ncols = no;
    McicaArgs m{};
    m.geo = G; m.nlay = nlay; m.nz = nz; m.ngpt = NGLW; m.permuteseed = 150; m.ncols = no; m.W = a.ws.W; m.icloud = in->icloud;
    m.cap = (int)cap; m.col0 = o0; m.lw_buffer = 1; m.cols = a.ws.cols; m.p3d = a.p3d; m.p8w = a.p8w; m.cldfra3d = a.cf.cldfra3d;
    m.mask = a.ws.mask; m.anyc = a.ws.anyc;
    cudaStream_t so = o0 == 0 ? spl : g.stream;
    { Timed t("lw_mcica", so); launch_mcica(m, so); }
    { Timed t("lw_prep", so); launch_lw_prep(a, so); }
    if (a.dbg.cldmask) {
      k_unpack_mask<<<(no + 127) / 128, 128, 0, so>>>(a.ws.mask, a.ws.cols, o0, no, (int)cap, a.ws.W, nlay, NGLW, a.dbg.cldmask);
      count_launch();
    }
    if (so != g.stream) { CK(cudaEventRecord(g.ev_pre_lw, so)); CK(cudaStreamWaitEvent(g.stream, g.ev_pre_lw, 0)); }
    cudaStream_t s2 = g.overlap ? g.stream2 : g.stream;
    int kc = 0;
    for (int c0 = 0; c0 < no; c0 += (int)pcap, kc++) {
      LwArgs b = a;
      b.ncols = std::min((int)pcap, no - c0);
      b.col0 = o0 + c0;
      if (b.ws.cols) b.ws.cols += c0;
      b.ws.coef += (size_t)c0 * LWC_N; b.ws.aer += c0; b.ws.cld += c0; b.ws.mask += c0; b.ws.anyc += c0; b.ws.laytrop += c0; b.ws.colf += c0;
      b.ws.secdiff += c0;
      const int buf = g.overlap ? (kc & 1) : 0;
      b.ws.bpart += (size_t)buf * lw_sweep_groups() * (nlay + 1) * a.ws.nk * pcap;
      if (g.overlap && kc >= 2) CK(cudaStreamWaitEvent(g.stream, g.ev_swept[buf], 0));      // partials of chunk k-2 consumed
      { Timed t("lw_solve"); launch_lw_band(b, g.stream); }
      if (g.overlap) { CK(cudaEventRecord(g.ev_solved, g.stream)); CK(cudaStreamWaitEvent(s2, g.ev_solved, 0)); }
      { Timed t("lw_reduce", s2); launch_lw_reduce(b, s2); }
      if (g.overlap) CK(cudaEventRecord(g.ev_swept[buf], s2));
    }
    // (the LW call of a chained pair leaves its last sweeps running: the SW call joins them)
    const bool last_outer = o0 + (int)cap >= ncol;
    if (g.overlap && !(g.chain == 1 && last_outer)) {
      CK(cudaStreamWaitEvent(g.stream, g.ev_swept[0], 0)); CK(cudaStreamWaitEvent(g.stream, g.ev_swept[1], 0));
    }
  }
  if (g.chain == 1) { if (g.async_pair) CK(cudaEventRecord(g.ev_lw_done, g.stream2)); return 0; }
  return finish_call(dbglist);
}

int arc_rad_lw(const ArcDims *d, const ArcLwIn *in, ArcLwOut *out) { return arc_rad_lw_debug(d, in, out, nullptr); }

// One radiation step = RRTMG_LWRAD then RRTMG_SWRAD on the same tile (the order radiation_driver uses, DRV:1526-2009).
// With host arrays both run inside ONE slab pipeline: the arrays the two adapters share are uploaded once and the
// pipeline fills and drains once; otherwise this is simply the two calls.
int arc_rad_lwsw(const ArcDims *d, const ArcLwIn *lwin, ArcLwOut *lwout, const ArcSwIn *swin, ArcSwOut *swout) {
  API_LOCK;
  if (!g.ready) { g.err = "arc_rad_lwsw: not initialised"; return ARC_ERR_NOT_INIT; }
  if (!d || !lwin || !lwout || !swin || !swout) { g.err = "arc_rad_lwsw: null argument"; return ARC_ERR_BAD_ARG; }
  int rc = check_dims(*d);
  if (rc) return rc;
  if (lwin->memspace == ARC_MEM_HOST && swin->memspace == ARC_MEM_HOST && !swin->tauaer3d_sw && swin->aer_opt != 1) {
    const PipePart parts[2] = {lw_part(lwin, lwout), sw_part(swin, swout)};
    rc = run_pipelined(*d, parts, 2);
    if (rc != -1) return rc;
  }
  if (g.overlap && lwin->memspace == ARC_MEM_DEVICE && swin->memspace == ARC_MEM_DEVICE) {
    // one continuous pipeline: the SW column kernels run beside the LW ones, the SW solver follows the LW solver on the main
    // stream while the last LW sweep is still running; one join + synchronisation at the end
    CK(cudaEventRecord(g.ev_entry, g.stream)); CK(cudaStreamWaitEvent(g.stream3, g.ev_entry, 0));
    g.chain = 1;
    rc = arc_rad_lw(d, lwin, lwout);
    if (rc) { g.chain = 0; cudaStreamSynchronize(g.stream2); cudaStreamSynchronize(g.stream); collect_times(); return rc; }
    g.chain = 2;
    rc = arc_rad_sw(d, swin, swout);
    g.chain = 0;
    if (rc) { cudaStreamSynchronize(g.stream3); cudaStreamSynchronize(g.stream2); cudaStreamSynchronize(g.stream); }
    return rc;
  }
  if ((rc = arc_rad_lw(d, lwin, lwout))) return rc;
  return arc_rad_sw(d, swin, swout);
}

// ---------------------------------------------------------------------------------------------------------
// radiation_driver bookkeeping around the two calls (module_radiation_driver.F:1692-1702, 2180-2194)
__global__ void k_driver_post(Geo G, const float *__restrict__ lw, const float *__restrict__ sw, float *__restrict__ rthraten,
                              const float *__restrict__ gsw, const float *__restrict__ albedo, float *__restrict__ swdown) {
  const int tc = blockIdx.x * blockDim.x + threadIdx.x;
  if (tc >= G.ncol_tile) return;
  int i, j; G.ij(tc, i, j);
  if (rthraten)
    for (int k = G.kts; k <= G.kte; k++) { const size_t q = G.at3(i, k, j); rthraten[q] = lw[q] + sw[q]; }
  if (swdown) { const size_t ij = G.at2(i, j); swdown[ij] = gsw[ij] / (1.f - albedo[ij]); }
}

int arc_rad_driver_post(const ArcDims *d, int memspace, const float *rthratenlw, const float *rthratensw, float *rthraten,
                        const float *gsw, const float *albedo, float *swdown) {
  API_LOCK;
  if (!g.ready) { g.err = "arc_rad_driver_post: not initialised"; return ARC_ERR_NOT_INIT; }
  int rc = check_dims(*d);
  if (rc) return rc;
  if ((rthraten && !(rthratenlw && rthratensw)) || (swdown && !(gsw && albedo))) { g.err = "arc_rad_driver_post: missing input"; return ARC_ERR_BAD_ARG; }
  CK(cudaSetDevice(g.device));
  g.pool_next = 0; g.backs.clear();
  Geo G = make_geo(*d);
  const float *lw, *sw, *gs, *al; float *rt, *sd;
  if ((rc = in_arr(memspace, rthratenlw, G.n3(), &lw))) return rc;
  if ((rc = in_arr(memspace, rthratensw, G.n3(), &sw))) return rc;
  if ((rc = in_arr(memspace, gsw, G.n2(), &gs))) return rc;
  if ((rc = in_arr(memspace, albedo, G.n2(), &al))) return rc;
  if ((rc = out_arr(memspace, rthraten, G.n3(), &rt))) return rc;
  if ((rc = out_arr(memspace, swdown, G.n2(), &sd))) return rc;
  k_driver_post<<<(G.ncol_tile + 255) / 256, 256, 0, g.stream>>>(G, lw, sw, rt, gs, al, sd);
  count_launch();
  if ((rc = copy_back())) return rc;
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// radconst / calc_coszen (module_radiation_driver.F:2595-2666) and the flux accumulation (DRV:2308-2377)
void arc_rad_radconst(float xtime, float julian, float degrad, float dpd, float *declin, float *solcon) {
  (void)xtime;
  const float obecl = 23.5f * degrad;
  const float sinob = sinf(obecl);
  float sxlong;
  if (julian >= 80.f) sxlong = dpd * (julian - 80.f); else sxlong = dpd * (julian + 285.f);
  sxlong = sxlong * degrad;
  const float arg = sinob * sinf(sxlong);
  if (declin) *declin = asinf(arg);
  const float djul = julian * 360.f / 365.f;
  const float rjul = djul * degrad;
  const float eccfac = 1.000110f + 0.034221f * cosf(rjul) + 0.001280f * sinf(rjul) + 0.000719f * cosf(2 * rjul) + 0.000077f * sinf(2 * rjul);
  if (solcon) *solcon = 1370.f * eccfac;
}

__global__ void k_calc_coszen(Geo G, float xt24, float gmt, float declin, float degrad, const float *__restrict__ xlon,
                              const float *__restrict__ xlat, float *__restrict__ coszen, float *__restrict__ hrang) {
  const int tc = blockIdx.x * blockDim.x + threadIdx.x;
  if (tc >= G.ncol_tile) return;
  int i, j; G.ij(tc, i, j);
  const size_t q = G.at2(i, j);
  const float tloctm = gmt + xt24 / 60.f + xlon[q] / 15.f;
  const float hr = 15.f * (tloctm - 12.f) * degrad;
  const float xxlat = xlat[q] * degrad;
  if (hrang) hrang[q] = hr;
  coszen[q] = sinf(xxlat) * sinf(declin) + cosf(xxlat) * cosf(declin) * cosf(hr);
}

int arc_rad_calc_coszen(const ArcDims *d, int memspace, float julian, float xtime, float gmt, float declin, float degrad,
                        const float *xlon, const float *xlat, float *coszen, float *hrang) {
  API_LOCK;
  if (!g.ready) { g.err = "arc_rad_calc_coszen: not initialised"; return ARC_ERR_NOT_INIT; }
  if (!d || !xlon || !xlat || !coszen) { g.err = "arc_rad_calc_coszen: null argument"; return ARC_ERR_BAD_ARG; }
  int rc = check_dims(*d);
  if (rc) return rc;
  CK(cudaSetDevice(g.device));
  g.pool_next = 0; g.backs.clear();
  Geo G = make_geo(*d);
  // equation of time (jararias 2013), scalar part evaluated on the host in FP32 like the reference
  const float da = 6.2831853071795862f * (julian - 1) / 365.f;
  const float eot = (0.000075f + 0.001868f * cosf(da) - 0.032077f * sinf(da) - 0.014615f * cosf(2 * da) - 0.04089f * sinf(2 * da)) * (229.18f);
  const float xt24 = fmodf(xtime, 1440.f) + eot;
  const float *lo, *la; float *cz, *hr;
  if ((rc = in_arr(memspace, xlon, G.n2(), &lo))) return rc;
  if ((rc = in_arr(memspace, xlat, G.n2(), &la))) return rc;
  if ((rc = out_arr(memspace, coszen, G.n2(), &cz))) return rc;
  if ((rc = out_arr(memspace, hrang, G.n2(), &hr))) return rc;
  k_calc_coszen<<<(G.ncol_tile + 255) / 256, 256, 0, g.stream>>>(G, xt24, gmt, declin, degrad, lo, la, cz, hr);
  count_launch();
  if ((rc = copy_back())) return rc;
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaGetLastError());
  return 0;
}

__global__ void k_accumulate(Geo G, float dt, int nf, const float *const *__restrict__ flux, float *const *__restrict__ acc) {
  const int tc = blockIdx.x * blockDim.x + threadIdx.x;
  if (tc >= G.ncol_tile) return;
  int i, j; G.ij(tc, i, j);
  const size_t q = G.at2(i, j);
  for (int f = 0; f < nf; f++) acc[f][q] = __fadd_rn(acc[f][q], __fmul_rn(flux[f][q], dt));   // unfused like the reference
}

int arc_rad_accumulate(const ArcDims *d, int memspace, float dtaccum, int nfields, const float *const *flux, float *const *acc) {
  API_LOCK;
  if (!g.ready) { g.err = "arc_rad_accumulate: not initialised"; return ARC_ERR_NOT_INIT; }
  if (!d || !flux || !acc || nfields < 1 || nfields > 32) { g.err = "arc_rad_accumulate: bad argument"; return ARC_ERR_BAD_ARG; }
  int rc = check_dims(*d);
  if (rc) return rc;
  CK(cudaSetDevice(g.device));
  g.pool_next = 0; g.backs.clear();
  Geo G = make_geo(*d);
  const float *df[32]; float *da[32];
  for (int f = 0; f < nfields; f++) {
    if (!flux[f] || !acc[f]) { g.err = "arc_rad_accumulate: null field"; return ARC_ERR_BAD_ARG; }
    if ((rc = in_arr(memspace, flux[f], G.n2(), &df[f]))) return rc;
    if ((rc = out_arr(memspace, acc[f], G.n2(), &da[f]))) return rc;
  }
  void *pf, *pa;
  if ((rc = stage_slot(sizeof(float *) * 32, &pf))) return rc;
  if ((rc = stage_slot(sizeof(float *) * 32, &pa))) return rc;
  CK(cudaMemcpyAsync(pf, df, sizeof(float *) * nfields, cudaMemcpyHostToDevice, g.stream));
  CK(cudaMemcpyAsync(pa, da, sizeof(float *) * nfields, cudaMemcpyHostToDevice, g.stream));
  k_accumulate<<<(G.ncol_tile + 255) / 256, 256, 0, g.stream>>>(G, dtaccum, nfields, (const float *const *)pf, (float *const *)pa);
  count_launch();
  if ((rc = copy_back())) return rc;
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Domain statistics of 2-D diagnostic fields over the tile: [sum, sum of squares, count, min, max] per field, the
// quantities the offline decomposition needs for means / SD / SE (analysis_scripts/NCL_extraction_package/
// misc_stats_library.ncl:396-461, RadDecomp_functions.py:119-129).  One block per field, fixed-order tree: reproducible.
__global__ void __launch_bounds__(1024) k_domain_stats(Geo G, int nfields, const float *const *__restrict__ fields, double *__restrict__ out) {
  const int f = blockIdx.x;
  const float *x = fields[f];
  double s = 0.0, s2 = 0.0; float mn = INFINITY, mx = -INFINITY;
  for (int tc = threadIdx.x; tc < G.ncol_tile; tc += blockDim.x) {
    int i, j; G.ij(tc, i, j);
    const float v = x[G.at2(i, j)];
    s += (double)v; s2 += (double)v * (double)v; mn = fminf(mn, v); mx = fmaxf(mx, v);
  }
  __shared__ double sh[3][1024];
  __shared__ float shm[2][1024];
  sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = s2; shm[0][threadIdx.x] = mn; shm[1][threadIdx.x] = mx;
  __syncthreads();
  for (int off = 512; off > 0; off >>= 1) {
    if (threadIdx.x < off) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + off]; sh[1][threadIdx.x] += sh[1][threadIdx.x + off];
      shm[0][threadIdx.x] = fminf(shm[0][threadIdx.x], shm[0][threadIdx.x + off]);
      shm[1][threadIdx.x] = fmaxf(shm[1][threadIdx.x], shm[1][threadIdx.x + off]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[5 * f + 0] = sh[0][0]; out[5 * f + 1] = sh[1][0]; out[5 * f + 2] = (double)G.ncol_tile;
    out[5 * f + 3] = (double)shm[0][0]; out[5 * f + 4] = (double)shm[1][0];
  }
}

int arc_rad_domain_stats(const ArcDims *d, int memspace, int nfields, const float *const *fields, double *out) {
  API_LOCK;
  if (!g.ready) { g.err = "arc_rad_domain_stats: not initialised"; return ARC_ERR_NOT_INIT; }
  if (!d || !fields || !out || nfields < 1 || nfields > 64) { g.err = "arc_rad_domain_stats: bad argument"; return ARC_ERR_BAD_ARG; }
  int rc = check_dims(*d);
  if (rc) return rc;
  CK(cudaSetDevice(g.device));
  g.pool_next = 0; g.backs.clear();
  Geo G = make_geo(*d);
  const float *dev[64];
  for (int f = 0; f < nfields; f++) {
    if (!fields[f]) { g.err = "arc_rad_domain_stats: null field"; return ARC_ERR_BAD_ARG; }
    if ((rc = in_arr(memspace, fields[f], G.n2(), &dev[f]))) return rc;
  }
  void *dptrs; if ((rc = stage_slot(sizeof(float *) * 64, &dptrs))) return rc;
  CK(cudaMemcpyAsync(dptrs, dev, sizeof(float *) * nfields, cudaMemcpyHostToDevice, g.stream));
  double *dout = out;
  if (memspace != ARC_MEM_DEVICE) { void *p; if ((rc = stage_slot(sizeof(double) * 5 * 64, &p))) return rc; dout = (double *)p; }
  k_domain_stats<<<nfields, 1024, 0, g.stream>>>(G, nfields, (const float *const *)dptrs, dout);
  count_launch();
  if (memspace != ARC_MEM_DEVICE) CK(cudaMemcpyAsync(out, dout, sizeof(double) * 5 * nfields, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaGetLastError());
  return 0;
}

// Moran's I of 2-D diagnostic fields over the tile as calc_morans_i_2D computes it with the options calc_standard_stats
// passes (neighbour + manhattan; analysis_scripts/NCL_extraction_package/misc_stats_library.ncl:196-371, 401-404): the eight
// raveled-index displacements with wrap-around, weight 1 where the Manhattan distance of the two cells is exactly 1 -
// i.e. every ordered pair of edge-sharing cells once - deviations from the float mean, sums in double, normalised by
// W_sum * stddev^2 with the N-1 stddev.  calc_standard_stats multiplies the standard error by it (its "corrected SE").
// One block per field, fixed-order tree: reproducible.
// STAT_NB blocks per field, each over a contiguous range of cells; block partials are combined in block order: reproducible.
constexpr int STAT_NB = 32;
__device__ __forceinline__ double block_sum(double v, double *sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int off = 512; off > 0; off >>= 1) { if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off]; __syncthreads(); }
  const double r = sh[0];
  __syncthreads();
  return r;
}
__global__ void __launch_bounds__(1024) k_morans_sum(Geo G, const float *const *__restrict__ fields, double *__restrict__ part) {
  const int f = blockIdx.x, nb = gridDim.y, b = blockIdx.y;
  const float *x = fields[f];
  __shared__ double sh[1024];
  const int per = (G.ncol_tile + nb - 1) / nb, lo = b * per, hi = min(G.ncol_tile, lo + per);
  double s = 0.0;
  for (int tc = lo + threadIdx.x; tc < hi; tc += blockDim.x) { int i, j; G.ij(tc, i, j); s += (double)x[G.at2(i, j)]; }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) part[f * nb + b] = s;
}
__global__ void __launch_bounds__(1024) k_morans_pairs(Geo G, const float *const *__restrict__ fields, const double *__restrict__ part,
                                                        double *__restrict__ part2) {
  const int f = blockIdx.x, nb = gridDim.y, b = blockIdx.y;
  const float *x = fields[f];
  __shared__ double sh[1024];
  double tot = 0.0;
  for (int q = 0; q < nb; q++) tot += part[f * nb + q];
  const float mean = (float)(tot / (double)G.ncol_tile);
  const int per = (G.ncol_tile + nb - 1) / nb, lo = b * per, hi = min(G.ncol_tile, lo + per);
  // deviations are formed in single precision from the single-precision mean (X_diff = data - X_mean), products in double
  double ss = 0.0, au = 0.0;
  for (int tc = lo + threadIdx.x; tc < hi; tc += blockDim.x) {
    int i, j; G.ij(tc, i, j);
    const float d0 = x[G.at2(i, j)] - mean;
    ss += (double)d0 * (double)d0;
    double nbr = 0.0;
    if (i < G.ite) nbr += (double)(x[G.at2(i + 1, j)] - mean);
    if (j < G.jte) nbr += (double)(x[G.at2(i, j + 1)] - mean);
    au += 2.0 * (double)d0 * nbr;               // each edge-sharing pair counts in both directions
  }
  ss = block_sum(ss, sh);
  au = block_sum(au, sh);
  if (threadIdx.x == 0) { part2[(f * nb + b) * 2] = ss; part2[(f * nb + b) * 2 + 1] = au; }
}
__global__ void k_morans_final(Geo G, int nfields, int nb, const double *__restrict__ part2, float *__restrict__ out) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nfields) return;
  double ss = 0.0, au = 0.0;
  for (int q = 0; q < nb; q++) { ss += part2[(f * nb + q) * 2]; au += part2[(f * nb + q) * 2 + 1]; }
  const int ni = G.nci, nj = G.ncol_tile / G.nci;
  const double n = (double)G.ncol_tile;
  const float sd = (float)sqrt(ss / fmax(n - 1.0, 1.0));           // NCL stddev: N-1
  const double wsum = 2.0 * ((double)nj * (ni - 1) + (double)ni * (nj - 1));
  const float var = sd * sd;
  out[f] = (ss == 0.0 || wsum == 0.0) ? 0.f : (float)(au / (wsum * (double)var));
}

int arc_rad_morans_i(const ArcDims *d, int memspace, int nfields, const float *const *fields, float *out) {
  API_LOCK;
  if (!g.ready) { g.err = "arc_rad_morans_i: not initialised"; return ARC_ERR_NOT_INIT; }
  if (!d || !fields || !out || nfields < 1 || nfields > 64) { g.err = "arc_rad_morans_i: bad argument"; return ARC_ERR_BAD_ARG; }
  int rc = check_dims(*d);
  if (rc) return rc;
  CK(cudaSetDevice(g.device));
  g.pool_next = 0; g.backs.clear();
  Geo G = make_geo(*d);
  const float *dev[64];
  for (int f = 0; f < nfields; f++) {
    if (!fields[f]) { g.err = "arc_rad_morans_i: null field"; return ARC_ERR_BAD_ARG; }
    if ((rc = in_arr(memspace, fields[f], G.n2(), &dev[f]))) return rc;
  }
  void *dptrs; if ((rc = stage_slot(sizeof(float *) * 64, &dptrs))) return rc;
  CK(cudaMemcpyAsync(dptrs, dev, sizeof(float *) * nfields, cudaMemcpyHostToDevice, g.stream));
  float *dout = out;
  if (memspace != ARC_MEM_DEVICE) { void *p; if ((rc = stage_slot(sizeof(float) * 64, &p))) return rc; dout = (float *)p; }
  void *part; if ((rc = stage_slot(sizeof(double) * 64 * STAT_NB * 3, &part))) return rc;
  double *p1 = (double *)part, *p2 = p1 + 64 * STAT_NB;
  k_morans_sum<<<dim3(nfields, STAT_NB), 1024, 0, g.stream>>>(G, (const float *const *)dptrs, p1);
  k_morans_pairs<<<dim3(nfields, STAT_NB), 1024, 0, g.stream>>>(G, (const float *const *)dptrs, p1, p2);
  k_morans_final<<<1, 64, 0, g.stream>>>(G, nfields, STAT_NB, p2, dout);
  count_launch(3);
  if (memspace != ARC_MEM_DEVICE) CK(cudaMemcpyAsync(out, dout, sizeof(float) * nfields, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaGetLastError());
  return 0;
}

// cal_cldfra1 (module_radiation_driver.F:2886-3122, called for icloud = 1 at DRV:1104-1118): the cloud fraction radiation_driver
// computes before the RRTMG calls, here on the device so that CLDFRA need not be produced on the host.
int arc_rad_cal_cldfra1(const ArcDims *d, int memspace, const float *qv, const float *qc, const float *qi, const float *qs, int f_qv, int f_qc,
                        int f_qi, int f_qs, const float *t_phy, const float *p_phy, const float *f_ice_phy, int mp_physics, float *cldfra,
                        int *cldfra1_flag) {
  API_LOCK;
  if (!g.ready) { g.err = "arc_rad_cal_cldfra1: not initialised"; return ARC_ERR_NOT_INIT; }
  if (!d || !qv || !t_phy || !p_phy || !cldfra) { g.err = "arc_rad_cal_cldfra1: null argument"; return ARC_ERR_BAD_ARG; }
  int rc = check_dims(*d);
  if (rc) return rc;
  if ((f_qc > 0 && !qc) || (f_qi > 0 && !qi) || (f_qs > 0 && !qs)) { g.err = "arc_rad_cal_cldfra1: F_Qx set but array missing"; return ARC_ERR_BAD_ARG; }
  CK(cudaSetDevice(g.device));
  g.pool_next = 0; g.backs.clear();
  Geo G = make_geo(*d);
  const size_t n3 = G.n3();
  const float *dqv, *dqc = nullptr, *dqi = nullptr, *dqs = nullptr, *dt, *dp, *dfi = nullptr;
  if ((rc = in_arr(memspace, qv, n3, &dqv)) || (rc = in_arr(memspace, t_phy, n3, &dt)) || (rc = in_arr(memspace, p_phy, n3, &dp))) return rc;
  if (qc && (rc = in_arr(memspace, qc, n3, &dqc))) return rc;
  if (qi && (rc = in_arr(memspace, qi, n3, &dqi))) return rc;
  if (qs && (rc = in_arr(memspace, qs, n3, &dqs))) return rc;
  if (f_ice_phy && (rc = in_arr(memspace, f_ice_phy, n3, &dfi))) return rc;
  float *dcf; int *dfl = nullptr;
  if ((rc = out_arr(memspace, cldfra, n3, &dcf))) return rc;
  if (cldfra1_flag && (rc = out_arr(memspace, (float *)cldfra1_flag, n3, (float **)&dfl))) return rc;
  launch_cal_cldfra1(G, dqv, dqc, dqi, dqs, f_qv, f_qc, f_qi, f_qs, dt, dp, dfi, mp_physics, dcf, dfl, g.stream);
  if ((rc = copy_back())) return rc;
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaGetLastError());
  return 0;
}

// cal_cldfra2 (module_radiation_driver.F:2801-2874, called for icloud = 2 at DRV:1205): binary cloud fraction.
int arc_rad_cal_cldfra2(const ArcDims *d, int memspace, const float *qc, const float *qi, int f_qc, int f_qi, float *cldfra) {
  API_LOCK;
  if (!g.ready) { g.err = "arc_rad_cal_cldfra2: not initialised"; return ARC_ERR_NOT_INIT; }
  if (!d || !cldfra) { g.err = "arc_rad_cal_cldfra2: null argument"; return ARC_ERR_BAD_ARG; }
  int rc = check_dims(*d);
  if (rc) return rc;
  if ((f_qc && !qc) || (f_qc && f_qi && !qi)) { g.err = "arc_rad_cal_cldfra2: F_QC / F_QI set but array missing"; return ARC_ERR_BAD_ARG; }
  CK(cudaSetDevice(g.device));
  g.pool_next = 0; g.backs.clear();
  Geo G = make_geo(*d);
  const size_t n3 = G.n3();
  const float *dqc = nullptr, *dqi = nullptr;
  if (f_qc && (rc = in_arr(memspace, qc, n3, &dqc))) return rc;
  if (f_qc && f_qi && (rc = in_arr(memspace, qi, n3, &dqi))) return rc;
  float *dcf;
  if ((rc = out_arr(memspace, cldfra, n3, &dcf))) return rc;
  launch_cal_cldfra2(G, dqc, dqi, f_qc != 0, f_qi != 0, dcf, g.stream);
  if ((rc = copy_back())) return rc;
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaGetLastError());
  return 0;
}

// cal_cldfra3 (module_radiation_driver.F:3140-3274, called for icloud = 3 at DRV:1228): G. Thompson's cloud-fraction scheme.
// CLDFRA is written, QC and QI are INOUT (the scheme adds sub-grid condensate to the fractional layers it finds), QS is read.
int arc_rad_cal_cldfra3(const ArcDims *d, int memspace, float *cldfra, const float *qv, float *qc, float *qi, const float *qs, const float *p,
                        const float *t, const float *rho, const float *xland, float gridkm) {
  API_LOCK;
  if (!g.ready) { g.err = "arc_rad_cal_cldfra3: not initialised"; return ARC_ERR_NOT_INIT; }
  if (!d || !cldfra || !qv || !qc || !qi || !qs || !p || !t || !rho || !xland) {
    g.err = "Can not use icloud = 3 option, missing QC or QI field.";       // the reference's message (DRV:1236) covers the null case
    return ARC_ERR_BAD_ARG;
  }
  int rc = check_dims(*d);
  if (rc) return rc;
  if (d->kte - d->kts < 4) { g.err = "arc_rad_cal_cldfra3: needs at least 5 levels"; return ARC_ERR_BAD_ARG; }
  CK(cudaSetDevice(g.device));
  g.pool_next = 0; g.backs.clear();
  Geo G = make_geo(*d);
  const size_t n3 = G.n3(), n2 = G.n2();
  const float *dqv, *dqs, *dp, *dt, *drho, *dxl;
  float *dcf, *dqc, *dqi;
  if ((rc = in_arr(memspace, qv, n3, &dqv)) || (rc = in_arr(memspace, qs, n3, &dqs)) || (rc = in_arr(memspace, p, n3, &dp)) ||
      (rc = in_arr(memspace, t, n3, &dt)) || (rc = in_arr(memspace, rho, n3, &drho)) || (rc = in_arr(memspace, xland, n2, &dxl))) return rc;
  if ((rc = out_arr(memspace, cldfra, n3, &dcf)) || (rc = out_arr(memspace, qc, n3, &dqc)) || (rc = out_arr(memspace, qi, n3, &dqi))) return rc;
  void *w[3];
  for (int q = 0; q < 3; q++) if ((rc = stage_slot(n3 * 4, &w[q]))) return rc;
  launch_cal_cldfra3(G, dcf, dqv, dqc, dqi, dqs, dp, dt, drho, dxl, gridkm, (float *)w[0], (float *)w[1], (float *)w[2], g.stream);
  if ((rc = copy_back())) return rc;
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaGetLastError());
  return 0;
}

// The date arithmetic ozn_time_int and aer_time_int share (DRV:4023-4082 = DRV:4268-4327): the two mid-month days that bracket
// JULIAN + 1 and their linear weights, December - January wrapping; the reference's scalar code in single precision.
static void clim_time_weights(float julian, int &nm_out, int &np_out, float &fact1_out, float &fact2_out) {
  static const int date_oz[12] = {16, 45, 75, 105, 136, 166, 197, 228, 258, 289, 319, 350};
  const float daysperyear = 365.f;
  volatile float intjulian = julian + 1.0f;                  // offset by one day (volatile: every step rounded to single)
  int ijul = (int)intjulian;
  intjulian = intjulian - (float)ijul;
  ijul = ijul % 365;
  if (ijul == 0) ijul = 365;
  intjulian = intjulian + (float)ijul;
  int np1 = 1; bool found = false;
  for (int m = 1; m <= 12; m++)
    if ((float)date_oz[m - 1] > intjulian && !found) { np1 = m; found = true; }
  const float cdayozp = (float)date_oz[np1 - 1];
  float cdayozm; int np, nm;
  if (np1 > 1) { cdayozm = (float)date_oz[np1 - 2]; np = np1; nm = np - 1; }
  else { cdayozm = (float)date_oz[11]; np = np1; nm = 12; }
  volatile float deltat, fact1, fact2;
  if (np1 == 1) {                                            // December - January
    deltat = cdayozp + daysperyear - cdayozm;
    if (intjulian > cdayozp) { fact1 = (cdayozp + daysperyear - intjulian) / deltat; fact2 = (intjulian - cdayozm) / deltat; }
    else { fact1 = (cdayozp - intjulian) / deltat; fact2 = (intjulian + daysperyear - cdayozm) / deltat; }
  } else {
    deltat = cdayozp - cdayozm;
    fact1 = (cdayozp - intjulian) / deltat;
    fact2 = (intjulian - cdayozm) / deltat;
  }
  nm_out = nm; np_out = np; fact1_out = fact1; fact2_out = fact2;
}

// ozn_time_int (module_radiation_driver.F:3993-4098; o3input = 2, DRV:1250): the monthly CAM ozone climatology interpolated in
// time to the model day.  The date arithmetic (which two months, which weights) is the reference's scalar code in single
// precision, on the host; the blend of the two months runs on the device.
int arc_rad_ozn_time_int(const ArcDims *d, int memspace, int julday, float julian, int levsiz, int num_months, const float *ozmixm, float *ozmixt) {
  API_LOCK;
  (void)julday;                                              // as in the reference: only JULIAN is used
  if (!g.ready) { g.err = "arc_rad_ozn_time_int: not initialised"; return ARC_ERR_NOT_INIT; }
  if (!d || !ozmixm || !ozmixt) { g.err = "arc_rad_ozn_time_int: null argument"; return ARC_ERR_BAD_ARG; }
  if (levsiz < 2 || num_months < 12) { g.err = "arc_rad_ozn_time_int: needs levsiz >= 2 and the 12 monthly fields"; return ARC_ERR_BAD_ARG; }
  int rc = check_dims(*d);
  if (rc) return rc;
  int nm, np; float fact1, fact2;
  clim_time_weights(julian, nm, np, fact1, fact2);
  CK(cudaSetDevice(g.device));
  g.pool_next = 0; g.backs.clear();
  Geo G = make_geo(*d);
  const size_t nlev = (size_t)G.ni * (size_t)levsiz * (size_t)(G.jme - G.jms + 1);
  const float *m0, *m1;
  if ((rc = in_arr(memspace, ozmixm + nlev * (size_t)(nm - 1), nlev, &m0)) || (rc = in_arr(memspace, ozmixm + nlev * (size_t)(np - 1), nlev, &m1))) return rc;
  float *dt;
  if ((rc = out_arr(memspace, ozmixt, nlev, &dt))) return rc;
  launch_ozn_time_int(G, levsiz, m0, m1, fact1, fact2, dt, g.stream);
  if ((rc = copy_back())) return rc;
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaGetLastError());
  return 0;
}

// ozn_p_int (module_radiation_driver.F:4100-4234; DRV:1256): ozone on the data pressure levels `pin` (HOST array, Pa, top
// down, strictly increasing) interpolated to the model mid-level pressures p(i,k,j) -> o3vmr(i,k,j), the O3RAD that
// RRTMG_SWRAD / RRTMG_LWRAD read with o3input = 2.
int arc_rad_ozn_p_int(const ArcDims *d, int memspace, const float *p, const float *pin, int levsiz, const float *ozmixt, float *o3vmr) {
  API_LOCK;
  if (!g.ready) { g.err = "arc_rad_ozn_p_int: not initialised"; return ARC_ERR_NOT_INIT; }
  if (!d || !p || !pin || !ozmixt || !o3vmr) { g.err = "arc_rad_ozn_p_int: null argument"; return ARC_ERR_BAD_ARG; }
  if (levsiz < 2 || levsiz > ARC_OZN_MAXLEV) { g.err = "arc_rad_ozn_p_int: levsiz must be 2.." + std::to_string(ARC_OZN_MAXLEV); return ARC_ERR_BAD_ARG; }
  for (int k = 1; k < levsiz; k++)
    if (!(pin[k] > pin[k - 1])) { g.err = "OZN_P_INT: Bad ozone data: non-monotonicity suspected"; return ARC_ERR_BAD_ARG; }
  int rc = check_dims(*d);
  if (rc) return rc;
  if (d->kts != 1) { g.err = "arc_rad_ozn_p_int: kts must be 1 (the reference indexes its work arrays from 1)"; return ARC_ERR_UNSUPPORTED; }
  CK(cudaSetDevice(g.device));
  g.pool_next = 0; g.backs.clear();
  Geo G = make_geo(*d);
  const size_t nlev = (size_t)G.ni * (size_t)levsiz * (size_t)(G.jme - G.jms + 1);
  const float *dp, *dt; float *dv;
  if ((rc = in_arr(memspace, p, G.n3(), &dp)) || (rc = in_arr(memspace, ozmixt, nlev, &dt))) return rc;
  if ((rc = out_arr(memspace, o3vmr, G.n3(), &dv))) return rc;
  launch_ozn_p_int(G, levsiz, pin, dp, dt, dv, g.stream);
  if ((rc = copy_back())) return rc;
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaGetLastError());
  return 0;
}

// aer_time_int (module_radiation_driver.F:4236-4343; aer_opt = 1): the monthly Tegen aerosol climatology
// aerodm(ims:ime, levsiz, jms:jme, num_months, no_src) -> aerodt(ims:ime, levsiz, jms:jme, no_src), as ozn_time_int per aerosol type.
int arc_rad_aer_time_int(const ArcDims *d, int memspace, int julday, float julian, int levsiz, int num_months, int no_src, const float *aerodm,
                         float *aerodt) {
  API_LOCK;
  (void)julday;
  if (!g.ready) { g.err = "arc_rad_aer_time_int: not initialised"; return ARC_ERR_NOT_INIT; }
  if (!d || !aerodm || !aerodt) { g.err = "arc_rad_aer_time_int: null argument"; return ARC_ERR_BAD_ARG; }
  if (levsiz < 2 || num_months < 12 || no_src < 1) { g.err = "arc_rad_aer_time_int: needs levsiz >= 2, 12 monthly fields, no_src >= 1"; return ARC_ERR_BAD_ARG; }
  int rc = check_dims(*d);
  if (rc) return rc;
  int nm, np; float fact1, fact2;
  clim_time_weights(julian, nm, np, fact1, fact2);
  CK(cudaSetDevice(g.device));
  g.pool_next = 0; g.backs.clear();
  Geo G = make_geo(*d);
  const size_t nlev = (size_t)G.ni * (size_t)levsiz * (size_t)(G.jme - G.jms + 1);
  for (int s = 0; s < no_src; s++) {
    const float *m0, *m1; float *dt;
    const float *base = aerodm + nlev * (size_t)num_months * (size_t)s;
    if ((rc = in_arr(memspace, base + nlev * (size_t)(nm - 1), nlev, &m0)) || (rc = in_arr(memspace, base + nlev * (size_t)(np - 1), nlev, &m1))) return rc;
    if ((rc = out_arr(memspace, aerodt + nlev * (size_t)s, nlev, &dt))) return rc;
    launch_ozn_time_int(G, levsiz, m0, m1, fact1, fact2, dt, g.stream);
  }
  if ((rc = copy_back())) return rc;
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaGetLastError());
  return 0;
}

// aer_p_int (module_radiation_driver.F:4345-4506): the climatology on its pressure levels `pin` (HOST array, hPa, top down) ->
// layer optical depths AEROD(i,k,j,1:no_src) at the model pressures p (Pa; compared as p * 0.01), each multiplied by the layer's
// interface-pressure difference pf(k) - pf(k+1), and their column total TOTAOD(i,j).  AEROD is what RRTMG_SWRAD reads with aer_opt = 1.
int arc_rad_aer_p_int(const ArcDims *d, int memspace, const float *p, const float *pin, int levsiz, const float *aerodt, float *aerod, int no_src,
                      const float *pf, float *totaod) {
  API_LOCK;
  if (!g.ready) { g.err = "arc_rad_aer_p_int: not initialised"; return ARC_ERR_NOT_INIT; }
  if (!d || !p || !pin || !aerodt || !aerod || !pf || !totaod) { g.err = "arc_rad_aer_p_int: null argument"; return ARC_ERR_BAD_ARG; }
  if (levsiz < 2 || levsiz > ARC_OZN_MAXLEV || no_src < 1) { g.err = "arc_rad_aer_p_int: levsiz must be 2.." + std::to_string(ARC_OZN_MAXLEV) + ", no_src >= 1"; return ARC_ERR_BAD_ARG; }
  for (int k = 1; k < levsiz; k++)
    if (!(pin[k] > pin[k - 1])) { g.err = "AER_P_INT: Bad aerosol data: non-monotonicity suspected"; return ARC_ERR_BAD_ARG; }
  int rc = check_dims(*d);
  if (rc) return rc;
  if (d->kts != 1) { g.err = "arc_rad_aer_p_int: kts must be 1 (the reference indexes its work arrays from 1)"; return ARC_ERR_UNSUPPORTED; }
  CK(cudaSetDevice(g.device));
  g.pool_next = 0; g.backs.clear();
  Geo G = make_geo(*d);
  const size_t nlev = (size_t)G.ni * (size_t)levsiz * (size_t)(G.jme - G.jms + 1);
  const float *dp, *dt, *dpf; float *dv, *dtot;
  if ((rc = in_arr(memspace, p, G.n3(), &dp)) || (rc = in_arr(memspace, pf, G.n3(), &dpf)) || (rc = in_arr(memspace, aerodt, nlev * no_src, &dt))) return rc;
  if ((rc = out_arr(memspace, aerod, G.n3() * no_src, &dv)) || (rc = out_arr(memspace, totaod, G.n2(), &dtot))) return rc;
  launch_clim_p_int(G, levsiz, pin, dp, 0.01f, no_src, dt, dv, dpf, dtot, g.stream);
  if ((rc = copy_back())) return rc;
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaGetLastError());
  return 0;
}

// Order statistics of 2-D diagnostic fields over the tile: calc_boxplot_stats picks sorted(x)[round(0.01 * p * (N - 1))] for
// p = 5, 25, 50, 75, 95 (misc_stats_library.ncl:145-189; calc_standard_stats stores them as median, quartiles and 5th / 95th
// percentile, ncl:439-445).  No sort: one block per (field, percentile) selects the element of that rank by four 8-bit radix
// passes over the order-preserving integer image of the floats (histogram in shared memory, digit by digit from the top), so
// the result is exactly the value the reference's qsort + index gives.
__device__ __forceinline__ uint32_t f2key(float x) { const uint32_t u = __float_as_uint(x); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float key2f(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }
// Selection state per (field, percentile): the key prefix fixed so far and the rank inside the cells that share it.  Each of
// the four passes is a histogram kernel (STAT_NB blocks per field; shared-memory histograms for all percentiles in ONE scan of
// the block's cells, added to the global histogram with integer atomics - order-independent, reproducible) and a selection kernel.
struct PercState { uint32_t prefix, rank; };
__global__ void __launch_bounds__(1024) k_perc_hist(Geo G, int nperc, int pass, const float *const *__restrict__ fields,
                                                     const PercState *__restrict__ st, unsigned *__restrict__ ghist) {
  const int f = blockIdx.x, nb = gridDim.y, b = blockIdx.y;
  const float *x = fields[f];
  __shared__ unsigned hist[16 * 256];
  __shared__ uint32_t pre[16];
  const int nh = pass == 0 ? 1 : nperc;                 // the first pass' histogram is common to all percentiles
  for (int t = threadIdx.x; t < nh * 256; t += blockDim.x) hist[t] = 0u;
  if (threadIdx.x < nperc) pre[threadIdx.x] = st[f * nperc + threadIdx.x].prefix;
  __syncthreads();
  const int shift = 24 - 8 * pass;
  const int per = (G.ncol_tile + nb - 1) / nb, lo = b * per, hi = min(G.ncol_tile, lo + per);
  for (int tc = lo + threadIdx.x; tc < hi; tc += blockDim.x) {
    int i, j; G.ij(tc, i, j);
    const uint32_t k = f2key(x[G.at2(i, j)]);
    const uint32_t dgt = (k >> shift) & 255u;
    if (pass == 0) atomicAdd(&hist[dgt], 1u);
    else {
      const uint32_t hi8 = k >> (shift + 8);
      for (int q = 0; q < nperc; q++) if (hi8 == pre[q]) atomicAdd(&hist[q * 256 + dgt], 1u);
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < nh * 256; t += blockDim.x) if (hist[t]) atomicAdd(&ghist[(size_t)f * 16 * 256 + t], hist[t]);
}
__global__ void k_perc_select(int nfields, int nperc, int pass, PercState *__restrict__ st, unsigned *__restrict__ ghist, float *__restrict__ out) {
  const int f = blockIdx.x, q = threadIdx.x;
  if (q < nperc) {
    const unsigned *h = ghist + (size_t)f * 16 * 256 + (pass == 0 ? 0 : q * 256);
    PercState s = st[f * nperc + q];
    unsigned cum = 0u; int dsel = 255;
    for (int dgt = 0; dgt < 256; dgt++) { if (s.rank < cum + h[dgt]) { dsel = dgt; break; } cum += h[dgt]; }
    s.rank -= cum; s.prefix = (s.prefix << 8) | (uint32_t)dsel;
    st[f * nperc + q] = s;
    if (pass == 3) out[f * nperc + q] = key2f(s.prefix);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 16 * 256; t += blockDim.x) ghist[(size_t)f * 16 * 256 + t] = 0u;     // ready for the next pass
}

int arc_rad_percentiles(const ArcDims *d, int memspace, int nfields, const float *const *fields, int nperc, const float *perc, float *out) {
  API_LOCK;
  if (!g.ready) { g.err = "arc_rad_percentiles: not initialised"; return ARC_ERR_NOT_INIT; }
  if (!d || !fields || !out || !perc || nfields < 1 || nfields > 64 || nperc < 1 || nperc > 16) { g.err = "arc_rad_percentiles: bad argument"; return ARC_ERR_BAD_ARG; }
  int rc = check_dims(*d);
  if (rc) return rc;
  CK(cudaSetDevice(g.device));
  g.pool_next = 0; g.backs.clear();
  Geo G = make_geo(*d);
  int ranks[16];
  for (int q = 0; q < nperc; q++) {
    if (!(perc[q] >= 0.f && perc[q] <= 100.f)) { g.err = "arc_rad_percentiles: percentile outside 0..100"; return ARC_ERR_BAD_ARG; }
    // pt_x = round(.01 * perc_point * (numel - 1), 3): single-precision product left to right, round half away from zero (ncl:176)
    const float ptx = 0.01f * perc[q] * (float)(G.ncol_tile - 1);
    long r = (long)floorf(ptx + 0.5f);
    ranks[q] = (int)std::min<long>(std::max<long>(r, 0), G.ncol_tile - 1);
  }
  const float *dev[64];
  for (int f = 0; f < nfields; f++) {
    if (!fields[f]) { g.err = "arc_rad_percentiles: null field"; return ARC_ERR_BAD_ARG; }
    if ((rc = in_arr(memspace, fields[f], G.n2(), &dev[f]))) return rc;
  }
  void *dptrs; if ((rc = stage_slot(sizeof(float *) * 64, &dptrs))) return rc;
  CK(cudaMemcpyAsync(dptrs, dev, sizeof(float *) * nfields, cudaMemcpyHostToDevice, g.stream));
  std::vector<PercState> hst((size_t)nfields * nperc);
  for (int f = 0; f < nfields; f++) for (int q = 0; q < nperc; q++) hst[(size_t)f * nperc + q] = PercState{0u, (uint32_t)ranks[q]};
  void *dst; if ((rc = stage_slot(sizeof(PercState) * 64 * 16, &dst))) return rc;
  CK(cudaMemcpyAsync(dst, hst.data(), sizeof(PercState) * hst.size(), cudaMemcpyHostToDevice, g.stream));
  void *dh; if ((rc = stage_slot(sizeof(unsigned) * 64 * 16 * 256, &dh))) return rc;
  CK(cudaMemsetAsync(dh, 0, sizeof(unsigned) * (size_t)nfields * 16 * 256, g.stream));
  float *dout = out;
  if (memspace != ARC_MEM_DEVICE) { void *p; if ((rc = stage_slot(sizeof(float) * 64 * 16, &p))) return rc; dout = (float *)p; }
  for (int pass = 0; pass < 4; pass++) {
    k_perc_hist<<<dim3(nfields, STAT_NB), 1024, 0, g.stream>>>(G, nperc, pass, (const float *const *)dptrs, (const PercState *)dst, (unsigned *)dh);
    k_perc_select<<<nfields, 256, 0, g.stream>>>(nfields, nperc, pass, (PercState *)dst, (unsigned *)dh, dout);
  }
  count_launch(8);
  if (memspace != ARC_MEM_DEVICE) CK(cudaMemcpyAsync(out, dout, sizeof(float) * nfields * nperc, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaGetLastError());
  return 0;
}

// Host-only probe used by the CPU test-suite: parse + reduce the k-distribution files exactly as arc_rad_init does
// (no GPU needed) and return one reduced table ("sw16.absa", "lw3.ka_mn2o", ...) or, with name == NULL, only validate.
int arc_rad_host_table(const char *inline_tables, const char *sw_data_path, const char *lw_data_path, float cp, float p_top,
                       int kme, const char *name, float *buf, int cap) {
  static HostTables T; static std::string key;
  if (!inline_tables || !sw_data_path || !lw_data_path) { g.err = "arc_rad_host_table: null argument"; return -ARC_ERR_BAD_ARG; }
  std::string k = std::string(inline_tables) + "|" + sw_data_path + "|" + lw_data_path;
  if (k != key) {
    int rc = build_host_tables(inline_tables, sw_data_path, lw_data_path, cp, p_top, kme, T, g.err);
    if (rc) { key.clear(); return -rc; }
    key = k;
  }
  if (!name) return 0;
  if (std::string(name) == "lw_nlayers") return T.lw_nlayers;
  auto it = T.reduced.find(name);
  if (it == T.reduced.end()) { g.err = std::string("no such table: ") + name; return -ARC_ERR_BAD_ARG; }
  const int n = (int)it->second.size();
  if (buf && cap >= n) memcpy(buf, it->second.data(), (size_t)n * 4);
  return n;
}

// Self-test (GPU): jp | jt << 8 | jt1 << 12 of setcoef for n host (p [hPa], T [K]) pairs, through the prep kernels' own code
int arc_rad_selftest_pt(const float *p, const float *t, int n, int *packed) {
  API_LOCK;
  if (!g.ready) { g.err = "arc_rad_selftest_pt: not initialised"; return ARC_ERR_NOT_INIT; }
  CK(cudaSetDevice(g.device));
  g.pool_next = 0; g.backs.clear();
  const float *dp, *dt; void *dout;
  int rc;
  if ((rc = in_arr(ARC_MEM_HOST, p, n, &dp))) return rc;
  if ((rc = in_arr(ARC_MEM_HOST, t, n, &dt))) return rc;
  if ((rc = stage_slot((size_t)n * 4, &dout))) return rc;
  launch_selftest_pt(g.D, dp, dt, n, (int *)dout, g.stream);
  CK(cudaMemcpyAsync(packed, dout, (size_t)n * 4, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaGetLastError());
  return 0;
}

// Self-test of glibc_math.cuh: which = 0 logf(x), 1 expf(x), 2 powf(x, y); on_device = 0 evaluates the host instantiation
// (no GPU, no init needed), 1 the device instantiation.  The caller compares with the C library.
__global__ void k_selftest_libm(int which, const float *__restrict__ x, const float *__restrict__ y, int n, float *__restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  out[t] = which == 0 ? glm::logf_(x[t]) : which == 1 ? glm::expf_(x[t]) : glm::powf_(x[t], y[t]);
}
int arc_rad_selftest_libm(int which, const float *x, const float *y, int n, float *out, int on_device) {
  API_LOCK;
  if (which < 0 || which > 2 || !x || !out || (which == 2 && !y) || n < 0) { g.err = "arc_rad_selftest_libm: bad argument"; return ARC_ERR_BAD_ARG; }
  if (!on_device) {
    for (int t = 0; t < n; t++) out[t] = which == 0 ? glm::logf_(x[t]) : which == 1 ? glm::expf_(x[t]) : glm::powf_(x[t], y[t]);
    return 0;
  }
  if (!g.ready) { g.err = "arc_rad_selftest_libm: not initialised"; return ARC_ERR_NOT_INIT; }
  CK(cudaSetDevice(g.device));
  g.pool_next = 0; g.backs.clear();
  const float *dx, *dy = nullptr; void *dout;
  int rc;
  if ((rc = in_arr(ARC_MEM_HOST, x, n, &dx))) return rc;
  if (which == 2 && (rc = in_arr(ARC_MEM_HOST, y, n, &dy))) return rc;
  if ((rc = stage_slot((size_t)n * 4, &dout))) return rc;
  k_selftest_libm<<<(n + 255) / 256, 256, 0, g.stream>>>(which, dx, dy, n, (float *)dout);
  count_launch();
  CK(cudaMemcpyAsync(out, dout, (size_t)n * 4, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaGetLastError());
  return 0;
}

// Self-test of the branch-free division used by the solver kernels against the compiler's IEEE division, over n
// pseudo-random operand pairs with magnitudes 1e-30 .. 1e+10 (and exact zeros as numerators).  Returns mismatches.
__global__ void k_selftest_div(int n, unsigned seed, int *bad) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  unsigned x = seed ^ (t * 2654435761u);
  auto rnd = [&]() { x ^= x << 13; x ^= x >> 17; x ^= x << 5; return x; };
  const float ea = -30.f + 40.f * (rnd() >> 8) * (1.0f / 16777216.0f), eb = -30.f + 40.f * (rnd() >> 8) * (1.0f / 16777216.0f);
  float a = exp10f(ea) * (1.f + (rnd() >> 9) * (1.0f / 8388608.0f)), b = exp10f(eb) * (1.f + (rnd() >> 9) * (1.0f / 8388608.0f));
  if ((rnd() & 63u) == 0u) a = 0.f;
  if (rnd() & 1u) a = -a;
  const float q0 = __fdiv_rn(a, b), q1 = div_rn(a, b);
  const bool normal = q0 == 0.f || (fabsf(q0) > 1e-37f && fabsf(q0) < 1e37f);
  if (normal && !(q0 == q1)) atomicAdd(bad, 1);     // value comparison: -0/b gives +0 here, -0 in IEEE
}
// rcp_rn against __frcp_rn for every float whose bit pattern lies in [lo_bits, hi_bits] (both signs).  Returns mismatches.
__global__ void k_selftest_rcp(unsigned lo_bits, unsigned count, unsigned long long *bad) {
  const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2ull * count) return;
  const unsigned bits = lo_bits + (unsigned)(t >> 1);
  const float x = __uint_as_float(bits | ((t & 1ull) ? 0x80000000u : 0u));
  if (!(__frcp_rn(x) == rcp_rn(x))) atomicAdd(bad, 1ull);
}
long long arc_rad_selftest_rcp(unsigned lo_bits, unsigned hi_bits) {
  API_LOCK;
  if (!g.ready || hi_bits < lo_bits) return -1;
  cudaSetDevice(g.device);
  unsigned long long *d; if (cudaMalloc(&d, 8) != cudaSuccess) return -1;
  cudaMemsetAsync(d, 0, 8, g.stream);
  const unsigned count = hi_bits - lo_bits + 1u;
  const unsigned long long nthreads = 2ull * count;
  k_selftest_rcp<<<(unsigned)((nthreads + 255) / 256), 256, 0, g.stream>>>(lo_bits, count, d);
  count_launch();
  unsigned long long h = ~0ull;
  cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, g.stream);
  cudaStreamSynchronize(g.stream);
  cudaFree(d);
  return (long long)h;
}
int arc_rad_selftest_div(int n, unsigned seed) {
  API_LOCK;
  if (!g.ready) return -1;
  cudaSetDevice(g.device);
  int *d; if (cudaMalloc(&d, 4) != cudaSuccess) return -1;
  cudaMemsetAsync(d, 0, 4, g.stream);
  k_selftest_div<<<(n + 255) / 256, 256, 0, g.stream>>>(n, seed, d);
  count_launch();
  int h = -1;
  cudaMemcpyAsync(&h, d, 4, cudaMemcpyDeviceToHost, g.stream);
  cudaStreamSynchronize(g.stream);
  cudaFree(d);
  return h;
}

// ---------------------------------------------------------------------------------------------------------
void arc_aer_default_refindex(float *refr, float *refi) {
  float a[AER_NCLASS][AER_NWL], b[AER_NCLASS][AER_NWL];
  default_refindex(a, b);
  if (refr) memcpy(refr, a, sizeof(a));
  if (refi) memcpy(refi, b, sizeof(b));
}

int arc_aer_init(const float *refr, const float *refi) {
  API_LOCK;
  if (!g.ready) { g.err = "arc_aer_init: call arc_rad_init first"; return ARC_ERR_NOT_INIT; }
  CK(cudaSetDevice(g.device));
  return aer_init(refr, refi, g.err);
}

int arc_aer_optics(const ArcDims *d, const ArcAerIn *in, ArcAerOut *out) {
  API_LOCK;
  if (!g.ready) { g.err = "arc_aer_optics: not initialised"; return ARC_ERR_NOT_INIT; }
  if (!d || !in || !out) { g.err = "arc_aer_optics: null argument"; return ARC_ERR_BAD_ARG; }
  int rc = check_dims(*d);
  if (rc) return rc;
  if (in->mode != ARC_AER_SECTIONAL && in->mode != ARC_AER_MODAL) { g.err = "arc_aer_optics: unknown mode"; return ARC_ERR_BAD_ARG; }
  if (in->nbin < 1 || in->nbin > ARC_AER_MAXBIN || !in->alt || !in->dz8w) { g.err = "arc_aer_optics: bad nbin / missing alt, dz8w"; return ARC_ERR_BAD_ARG; }
  CK(cudaSetDevice(g.device));
  if (!aer_ready() && (rc = aer_init(nullptr, nullptr, g.err))) return rc;
  g.pool_next = 0; g.backs.clear();
  const int ms = in->memspace;
  AerArgs a{};
  a.geo = make_geo(*d);
  const Geo &G = a.geo;
  const size_t n3 = G.n3();
  const int nz = d->kte - d->kts + 1;
  a.npts = G.ncol_tile * nz;
  a.nsec = in->mode == ARC_AER_SECTIONAL ? in->nbin : 8;
  AerSpecList sl{};
  sl.mode = in->mode; sl.nbin = in->nbin;
  for (int b = 0; b < in->nbin; b++) {
    if (in->nspec[b] < 1 || in->nspec[b] > ARC_AER_MAXSPEC || !in->num[b]) { g.err = "arc_aer_optics: bad species list"; return ARC_ERR_BAD_ARG; }
    sl.nspec[b] = in->nspec[b];
    sl.sigmag[b] = in->sigmag[b];
    if (in->mode == ARC_AER_MODAL && !(in->sigmag[b] > 1.0f)) { g.err = "arc_aer_optics: sigmag must be > 1 for modal input"; return ARC_ERR_BAD_ARG; }
    if ((rc = in_arr(ms, in->num[b], n3, &sl.num[b]))) return rc;
    for (int m = 0; m < in->nspec[b]; m++) {
      if (in->cls[b][m] < 0 || in->cls[b][m] >= ARC_CLS_N || !in->mass[b][m]) { g.err = "arc_aer_optics: bad species class / null mass array"; return ARC_ERR_BAD_ARG; }
      sl.cls[b][m] = in->cls[b][m];
      if ((rc = in_arr(ms, in->mass[b][m], n3, &sl.mass[b][m]))) return rc;
    }
  }
  if ((rc = in_arr(ms, in->alt, n3, &a.alt))) return rc;
  if ((rc = in_arr(ms, in->dz8w, n3, &a.dz8w))) return rc;
  for (int w = 0; w < 4; w++) {
    if (!out->tauaer[w] || !out->gaer[w] || !out->waer[w]) { g.err = "arc_aer_optics: SW output array missing"; return ARC_ERR_BAD_ARG; }
    if ((rc = out_arr(ms, out->tauaer[w], n3, &a.tauaer[w]))) return rc;
    if ((rc = out_arr(ms, out->gaer[w], n3, &a.gaer[w]))) return rc;
    if ((rc = out_arr(ms, out->waer[w], n3, &a.waer[w]))) return rc;
  }
  for (int w = 0; w < 16; w++) {
    if (!out->tauaerlw[w]) { g.err = "arc_aer_optics: LW output array missing"; return ARC_ERR_BAD_ARG; }
    if ((rc = out_arr(ms, out->tauaerlw[w], n3, &a.tauaerlw[w]))) return rc;
    if ((rc = out_arr(ms, out->extaerlw[w], n3, &a.extaerlw[w]))) return rc;
  }
  {
    Timed t("aer_optics");
    if ((rc = aer_run(a, sl, g.stream, g.err))) return rc;
  }
  if ((rc = copy_back())) return rc;
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaGetLastError());
  g.last_ms.erase("aer_optics");
  collect_times();
  return 0;
}

// CPU evaluation of the uploaded Chebyshev tables for one sphere (same interpolation as k_aer_mie) -- test tap
int arc_aer_table_eval(int wl, float radius_cm, float refr, float refi, float *qext, float *qsca, float *gg) {
  if (!aer_ready()) { g.err = "arc_aer_table_eval: tables not built"; return ARC_ERR_NOT_INIT; }
  const AerTables &T = aer_tables();
  if (wl < 0 || wl >= AER_NWL) return ARC_ERR_BAD_ARG;
  double r = std::min(std::max((double)radius_cm, T.rmin), T.rmax);
  double tr = (refr - T.refr_lo[wl]) / (T.refr_hi[wl] - T.refr_lo[wl]) * (AER_NREFR - 1);
  tr = std::min(std::max(tr, 0.0), (double)(AER_NREFR - 1));
  int ir = std::min((int)tr, AER_NREFR - 2); double t = tr - ir;
  double ti = (std::log(std::max((double)refi, 1e-30)) - std::log((double)T.refi_lo[wl])) / (std::log((double)T.refi_hi[wl]) - std::log((double)T.refi_lo[wl])) * (AER_NREFI - 1);
  ti = std::min(std::max(ti, 0.0), (double)(AER_NREFI - 1));
  int ii = std::min((int)ti, AER_NREFI - 2); double u = ti - ii;
  const double xrmin = std::log(T.rmin), xrmax = std::log(T.rmax);
  const double x = (2.0 * std::log(r) - xrmax - xrmin) / (xrmax - xrmin);
  double out[AER_NQ];
  for (int q = 0; q < AER_NQ; q++) {
    auto C = [&](int a_, int b_, int j) { return (double)T.coef[((((size_t)wl * AER_NQ + q) * AER_NREFR + a_) * AER_NREFI + b_) * AER_NCOEF_PAD + j]; };
    double tjm1 = 1.0, tj = x, acc = 0.0;
    for (int j = 0; j < AER_NCOEF; j++) {
      double v;
      if (j == 0) v = 0.5; else if (j == 1) v = x; else { v = 2.0 * x * tj - tjm1; tjm1 = tj; tj = v; }
      const double c = (1 - t) * (1 - u) * C(ir, ii, j) + t * (1 - u) * C(ir + 1, ii, j) + (1 - t) * u * C(ir, ii + 1, j) + t * u * C(ir + 1, ii + 1, j);
      acc += c * v;
    }
    out[q] = std::exp(acc);
  }
  *qext = (float)out[0]; *qsca = (float)std::min(out[1], out[0]); *gg = (float)out[2];
  return 0;
}
int arc_aer_mie_direct(int wl, float radius_cm, float refr, float refi, float *qext, float *qsca, float *gg) {
  if (wl < 0 || wl >= AER_NWL) return ARC_ERR_BAD_ARG;
  const double sw_um[4] = {0.30, 0.40, 0.60, 0.999};
  const double lw_nu[16] = {180., 425., 565., 665., 760., 900., 1030., 1130., 1285., 1435., 1640., 1940., 2165., 2315., 2490., 2925.};
  const double lam = wl < 4 ? sw_um[wl] * 1e-4 : 1.0 / lw_nu[wl - 4];
  double qe, qs, ga;
  mie_efficiencies(2.0 * M_PI * radius_cm / lam, refr, refi, qe, qs, ga);
  *qext = (float)qe; *qsca = (float)qs; *gg = (float)ga;
  return 0;
}

// FP32 FMA throughput of this GPU (TFLOP/s), measured with a dependent-chain-free FMA kernel: the roofline
// denominator for the FP32-pipe-bound solver kernels (MEASURED_PEAKS.json holds no FP32 figure).
__global__ void __launch_bounds__(256) k_fma_peak(float *out, int iters) {
  float a[16];
#pragma unroll
  for (int q = 0; q < 16; q++) a[q] = 1.0f + 1e-3f * (threadIdx.x + q);
  const float m = 0.999999f, c = 1e-7f;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int q = 0; q < 16; q++) a[q] = fmaf(a[q], m, c);
  }
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < 16; q++) s += a[q];
  if (s == 12345.678f) out[0] = s;
}
float arc_rad_measure_fp32_tflops(void) {
  API_LOCK;
  if (!g.ready) return -1.f;
  cudaSetDevice(g.device);
  float *dout; if (cudaMalloc(&dout, 4) != cudaSuccess) return -1.f;
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, g.device);
  const int blocks = pr.multiProcessorCount * 8, iters = 8192;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 0.f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0, g.stream);
    k_fma_peak<<<blocks, 256, 0, g.stream>>>(dout, iters);
    cudaEventRecord(e1, g.stream);
    cudaEventSynchronize(e1);
    float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 16.0 * iters * 256.0 * blocks;
    if (ms > 0.f) best = fmaxf(best, (float)(fl / (ms * 1e-3) / 1e12));
  }
  count_launch(5);
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(dout);
  return best;
}

}  // extern "C"
