// Host-side table preparation (see tables.h).  Schema-driven record parser + weight-matrix g-point
// reduction + per-(band, g) slice packing for TMA staging.
#include "tables.h"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

#include "../../include/arc_rad.h"

namespace arc {

const std::vector<float> &HostTables::get(const std::string &k) const {
  auto it = in.find(k);
  if (it == in.end()) { static std::vector<float> empty; fprintf(stderr, "arc_rad: inline table '%s' missing\n", k.c_str()); return empty; }
  return it->second;
}

namespace {

bool load_inline(const std::string &path, HostTables &T, std::string &err) {
  std::ifstream f(path, std::ios::binary);
  if (!f) { err = "cannot open inline table file " + path; return false; }
  std::vector<char> b((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  if (b.size() < 12 || memcmp(b.data(), "ARCTBL1\0", 8) != 0) { err = "bad magic in " + path; return false; }
  uint32_t n; memcpy(&n, &b[8], 4);
  for (uint32_t i = 0; i < n; i++) {
    const char *e = &b[12 + (size_t)i * 76];
    char name[33]; memcpy(name, e, 32); name[32] = 0;
    uint32_t nd, d[4]; uint64_t off;
    memcpy(&nd, e + 32, 4); memcpy(d, e + 36, 16); memcpy(&off, e + 68, 8);
    size_t cnt = (size_t)d[0] * d[1] * d[2] * d[3];
    if (off + cnt * 4 > b.size()) { err = "truncated inline table file"; return false; }
    std::vector<float> v(cnt); memcpy(v.data(), &b[off], cnt * 4);
    T.in[name] = std::move(v);
    T.in_dims[name] = {(int)d[0], (int)d[1], (int)d[2], (int)d[3]};
  }
  return true;
}

// ---- record schema ------------------------------------------------------------------------------
// item kinds:  s scalar float | i scalar int | v (16) | A kao | B kbo | m (rows,16) | g (16,cols) | t (ne,19,16)
struct Item { char kind; std::string name; int n; };

std::vector<Item> parse_schema(const char *s) {
  std::vector<Item> out;
  std::stringstream ss(s);
  std::string tok;
  while (ss >> tok) {
    Item it{tok[0], "", 0};
    size_t p1 = tok.find(':');
    if (p1 != std::string::npos) {
      size_t p2 = tok.find(':', p1 + 1);
      it.name = tok.substr(p1 + 1, p2 == std::string::npos ? std::string::npos : p2 - p1 - 1);
      if (p2 != std::string::npos) it.n = atoi(tok.c_str() + p2 + 1);
    }
    if (it.kind == 'A') it.name = "absa";
    if (it.kind == 'B') it.name = "absb";
    out.push_back(it);
  }
  return out;
}

const char *SW_SCHEMA[14] = {
    /*16*/ "s:rayl s:strrat i:layreffr A B m:selfref:10 m:forref:3 v:sflux",
    /*17*/ "s:rayl s:strrat i:layreffr A B m:selfref:10 m:forref:4 g:sflux:5",
    /*18*/ "s:rayl s:strrat i:layreffr A B m:selfref:10 m:forref:3 g:sflux:9",
    /*19*/ "s:rayl s:strrat i:layreffr A B m:selfref:10 m:forref:3 g:sflux:9",
    /*20*/ "s:rayl i:layreffr v:absch4 A B m:selfref:10 m:forref:4 v:sflux",
    /*21*/ "s:rayl s:strrat i:layreffr A B m:selfref:10 m:forref:4 g:sflux:9",
    /*22*/ "s:rayl s:strrat i:layreffr A B m:selfref:10 m:forref:3 g:sflux:9",
    /*23*/ "v:raylg s:givfac i:layreffr A m:selfref:10 m:forref:3 v:sflux",
    /*24*/ "g:rayla:9 v:raylb s:strrat i:layreffr v:abso3a v:abso3b A B m:selfref:10 m:forref:3 g:sflux:9",
    /*25*/ "v:raylg i:layreffr v:abso3a v:abso3b A v:sflux",
    /*26*/ "v:raylg v:sflux",
    /*27*/ "v:raylg s:scalekur i:layreffr A B v:sflux",
    /*28*/ "s:rayl s:strrat i:layreffr A B g:sflux:5",
    /*29*/ "s:rayl i:layreffr v:absh2o v:absco2 A B m:selfref:10 m:forref:4 v:sflux",
};
const char *LW_SF = " m:selfref:10 m:forref:4";
const char *LW_SCHEMA[16] = {
    /*1*/ "v:fracrefa v:fracrefb A B m:ka_mn2:19 m:kb_mn2:19",
    /*2*/ "v:fracrefa v:fracrefb A B",
    /*3*/ "g:fracrefa:9 g:fracrefb:5 A B t:ka_mn2o:9 t:kb_mn2o:5",
    /*4*/ "g:fracrefa:9 g:fracrefb:5 A B",
    /*5*/ "g:fracrefa:9 g:fracrefb:5 A B t:ka_mo3:9 v:ccl4",
    /*6*/ "v:fracrefa A m:ka_mco2:19 v:cfc11adj v:cfc12",
    /*7*/ "g:fracrefa:9 v:fracrefb A B t:ka_mco2:9 m:kb_mco2:19",
    /*8*/ "v:fracrefa v:fracrefb A B m:ka_mco2:19 m:kb_mco2:19 m:ka_mn2o:19 m:kb_mn2o:19 m:ka_mo3:19 v:cfc12 v:cfc22adj",
    /*9*/ "g:fracrefa:9 v:fracrefb A B t:ka_mn2o:9 m:kb_mn2o:19",
    /*10*/ "v:fracrefa v:fracrefb A B",
    /*11*/ "v:fracrefa v:fracrefb A B m:ka_mo2:19 m:kb_mo2:19",
    /*12*/ "g:fracrefa:9 A",
    /*13*/ "g:fracrefa:9 v:fracrefb A t:ka_mco2:9 t:ka_mco:9 m:kb_mo3:19",
    /*14*/ "v:fracrefa v:fracrefb A B",
    /*15*/ "g:fracrefa:9 A t:ka_mn2:9",
    /*16*/ "g:fracrefa:9 v:fracrefb A B",
};

struct Raw { std::vector<float> v; bool g_first = false; int lead = 1; };   // lead = elements per g (g last) or columns (g first)

// Fortran sequential unformatted file: every record is <int32 n> n bytes <int32 n>.  WRF is built with
// -fconvert=big-endian / -convert big_endian, so the RRTMG_*_DATA files it ships are big-endian; files written by a
// native little-endian program (ktables.py) are not.  The byte order is detected from the first record (the leading and
// trailing markers must agree and fit the file in exactly one of the two readings) and every 4-byte word
// (REAL*4 / INTEGER*4 payload, markers) is swapped when it is not the host's.
struct FileCursor {
  std::vector<char> buf; size_t pos = 0, rec_end = 0;
  uint32_t rec_len = 0;
  bool swap = false;
  static uint32_t bswap(uint32_t x) { return (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24); }
  uint32_t word(size_t at) const { uint32_t x; memcpy(&x, &buf[at], 4); return swap ? bswap(x) : x; }
  bool open(const std::string &p, std::string &err) {
    std::ifstream f(p, std::ios::binary);
    if (!f) { err = "cannot open " + p; return false; }
    buf.assign((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (buf.size() < 8) { err = p + ": not a Fortran unformatted file"; return false; }
    bool ok[2] = {false, false};
    for (int q = 0; q < 2; q++) {
      swap = q == 1;
      const uint64_t n = word(0);
      ok[q] = n > 0 && n % 4 == 0 && n + 8 <= buf.size() && word(4 + (size_t)n) == n;
    }
    if (ok[0] == ok[1]) { err = p + ": cannot determine the byte order of the record markers"; return false; }
    swap = ok[1];
    return true;
  }
  bool begin(std::string &err) {
    if (pos + 4 > buf.size()) { err = "unexpected end of file"; return false; }
    const int32_t n = (int32_t)word(pos); pos += 4; rec_end = pos + (size_t)n; rec_len = (uint32_t)n;
    if (n < 0 || rec_end + 4 > buf.size()) { err = "truncated record"; return false; }
    return true;
  }
  bool take(void *dst, size_t bytes, std::string &err) {
    if (pos + bytes > rec_end) { err = "record shorter than its schema"; return false; }
    memcpy(dst, &buf[pos], bytes); pos += bytes;
    if (swap) { uint32_t *w = (uint32_t *)dst; for (size_t i = 0; i < bytes / 4; i++) w[i] = bswap(w[i]); }
    return true;
  }
  bool end(std::string &err) {
    if (pos != rec_end) { err = "record longer than its schema"; return false; }
    if (word(pos) != rec_len) { err = "record markers disagree"; return false; }
    pos += 4;
    return true;
  }
};

// Reduction operator of one band as a dense (ngc x 16) weight matrix; row igc holds rwgt (or 1 for
// source terms) on the original g-points merged into igc, in ascending order.
struct Reducer {
  int ngc = 0;
  std::vector<int> first, count;          // original g range of each reduced point
  float rw[16];                           // rwgt of the band's 16 original points
  void apply(const Raw &r, bool weighted, std::vector<float> &out, int &lead_out) const {
    if (r.g_first) {
      int ncol = r.lead;
      out.assign((size_t)ngc * ncol, 0.f);
      for (int j = 0; j < ncol; j++)
        for (int q = 0; q < ngc; q++) {
          float s = 0.f;
          for (int p = first[q]; p < first[q] + count[q]; p++) {
            float x = r.v[(size_t)p + 16 * (size_t)j];
            s = s + (weighted ? x * rw[p] : x);
          }
          out[(size_t)q + (size_t)ngc * j] = s;
        }
      lead_out = ncol;
    } else {
      size_t lead = (size_t)r.lead;
      out.assign(lead * ngc, 0.f);
      for (size_t l = 0; l < lead; l++)
        for (int q = 0; q < ngc; q++) {
          float s = 0.f;
          for (int p = first[q]; p < first[q] + count[q]; p++) {
            float x = r.v[l + lead * (size_t)p];
            s = s + (weighted ? x * rw[p] : x);
          }
          out[l + lead * (size_t)q] = s;
        }
      lead_out = (int)lead;
    }
  }
};

std::vector<Reducer> make_reducers(int nb, const std::vector<float> &ngc, const std::vector<float> &ngn,
                                   const std::vector<float> &wt) {
  std::vector<Reducer> R(nb);
  size_t gcur = 0;
  for (int b = 0; b < nb; b++) {
    Reducer &r = R[b];
    r.ngc = (int)ngc[b];
    int p = 0;
    for (int q = 0; q < r.ngc; q++) {
      int c = (int)ngn[gcur + q];
      r.first.push_back(p); r.count.push_back(c); p += c;
    }
    gcur += r.ngc;
    for (int q = 0; q < r.ngc; q++) {
      float wsum = 0.f;
      for (int k = r.first[q]; k < r.first[q] + r.count[q]; k++) wsum = wsum + wt[k];
      for (int k = r.first[q]; k < r.first[q] + r.count[q]; k++) r.rw[k] = (r.ngc < 16) ? wt[k] / wsum : 1.0f;
    }
  }
  return R;
}

inline int pad4(int n) { return (n + 3) & ~3; }

struct BandRaw {
  std::map<std::string, Raw> arr;
  std::map<std::string, float> sc;
  std::map<std::string, int> ic;
};

bool read_band(FileCursor &F, const char *schema, int nspa, int nspb, BandRaw &B, std::string &err) {
  if (!F.begin(err)) return false;
  for (const Item &it : parse_schema(schema)) {
    Raw r;
    size_t cnt = 0;
    switch (it.kind) {
      case 's': { float x; if (!F.take(&x, 4, err)) return false; B.sc[it.name] = x; continue; }
      case 'i': { int32_t x; if (!F.take(&x, 4, err)) return false; B.ic[it.name] = x; continue; }
      case 'v': cnt = 16; r.lead = 1; break;
      case 'A': cnt = (size_t)std::max(nspa, 1) * 5 * 13 * 16; r.lead = (int)(cnt / 16); break;
      case 'B': cnt = (size_t)std::max(nspb, 1) * 5 * 47 * 16; r.lead = (int)(cnt / 16); break;
      case 'm': cnt = (size_t)it.n * 16; r.lead = it.n; break;
      case 't': cnt = (size_t)it.n * 19 * 16; r.lead = it.n * 19; break;
      case 'g': cnt = (size_t)16 * it.n; r.lead = it.n; r.g_first = true; break;
      default: err = "bad schema"; return false;
    }
    r.v.resize(cnt);
    if (!F.take(r.v.data(), cnt * 4, err)) return false;
    B.arr[it.name] = std::move(r);
  }
  return F.end(err);
}

void exp_tables(HostTables &T) {
  const int ntbl = 10000; const float pade = 0.278f, expeps = 1.e-20f;
  T.bpade = 1.0f / pade;
  T.sw_exp_tbl.assign(ntbl + 1, 0.f);
  T.lw_tau_tbl.assign(ntbl + 1, 0.f); T.lw_exp_tbl.assign(ntbl + 1, 0.f); T.lw_tfn_tbl.assign(ntbl + 1, 0.f);
  for (int i = 0; i <= ntbl; i++) {
    float tau, ex, tf;
    if (i == 0) { tau = 0.f; ex = 1.f; tf = 0.f; }
    else if (i == ntbl) { tau = 1.e10f; ex = expeps; tf = 1.f; }
    else {
      float tfn = (float)i / (float)ntbl;
      tau = T.bpade * tfn / (1.f - tfn);
      ex = expf(-tau);
      if (ex <= expeps) ex = expeps;
      tf = tau < 0.06f ? tau / 6.f : 1.f - 2.f * ((1.f / tau) - (ex / (1.f - ex)));
    }
    T.sw_exp_tbl[i] = ex; T.lw_tau_tbl[i] = tau; T.lw_exp_tbl[i] = ex; T.lw_tfn_tbl[i] = tf;
  }
}

}  // namespace

int build_host_tables(const std::string &inline_path, const std::string &sw_path, const std::string &lw_path,
                      float cp, float p_top, int kme, HostTables &T, std::string &err) {
  T = HostTables();
  if (!load_inline(inline_path, T, err)) return ARC_ERR_IO;
  T.cp = cp;
  const float grav = 9.8066f, secdy = 8.6400e4f;
  T.oneminus = 1.0f - 1.e-06f;
  T.pi = 2.0f * asinf(1.0f);
  T.heatfac = grav * secdy / (cp * 1.e2f);
  T.fluxfac = T.pi * 2.e4f;
  T.lw_nlayers = kme + (int)lroundf(p_top * 0.01f / 4.0f) - 1;     // LW:12861
  exp_tables(T);

  // ---------------- SW ----------------
  {
    auto R = make_reducers(14, T.get("sw_ngc"), T.get("sw_ngn"), T.get("sw_wt"));
    const auto &ngb = T.get("sw_ngb");
    for (int i = 0; i < 112; i++) T.sw_ngb[i] = (int)ngb[i] - 15;
    FileCursor F;
    if (!F.open(sw_path, err)) return ARC_ERR_IO;
    int g0 = 0;
    for (int ib = 0; ib < 14; ib++) {
      SwBandDesc &D = T.sw[ib];
      memset(&D, 0, sizeof(D));
      D.nspa = (int)T.get("sw_nspa")[ib]; D.nspb = (int)T.get("sw_nspb")[ib];
      D.ng = R[ib].ngc; D.g0 = g0; g0 += D.ng;
      BandRaw B;
      std::string e2;
      if (!read_band(F, SW_SCHEMA[ib], D.nspa, D.nspb, B, e2)) { err = "RRTMG_SW_DATA band " + std::to_string(ib + 16) + ": " + e2; return ARC_ERR_IO; }
      D.rayl = B.sc.count("rayl") ? B.sc["rayl"] : 0.f;
      D.strrat = B.sc.count("strrat") ? B.sc["strrat"] : 0.f;
      D.givfac = B.sc.count("givfac") ? B.sc["givfac"] : 1.f;
      D.scalekur = B.sc.count("scalekur") ? B.sc["scalekur"] : 1.f;
      D.layreffr = B.ic.count("layreffr") ? B.ic["layreffr"] : 0;
      std::map<std::string, std::vector<float>> red;
      std::map<std::string, int> lead;
      for (auto &kv : B.arr) {
        bool weighted = kv.first != "sflux";
        int l; R[ib].apply(kv.second, weighted, red[kv.first], l); lead[kv.first] = l;
        T.reduced["sw" + std::to_string(ib + 16) + "." + kv.first] = red[kv.first];
      }
      const int nA = red.count("absa") ? lead["absa"] : 0, nB = red.count("absb") ? lead["absb"] : 0;
      D.nfor = red.count("forref") ? lead["forref"] : 0;
      D.nsf = (B.arr["sflux"].g_first) ? lead["sflux"] : 1;
      int o = 0;
      D.oA = o; o += pad4(nA + 12);
      D.oB = o; o += pad4(nB + 12);
      D.oSelf = o; o += 12;
      D.oFor = o; o += 4;
      D.oSflx = o; o += 12;
      D.oRayl = o; o += 12;
      D.oMisc = o; o += 4;
      D.slice_floats = pad4(o);
      D.slice_base = (int)T.sw_buf.size();
      T.sw_buf.resize(T.sw_buf.size() + (size_t)D.slice_floats * D.ng, 0.f);
      for (int g = 0; g < D.ng; g++) {
        float *S = &T.sw_buf[(size_t)D.slice_base + (size_t)D.slice_floats * g];
        for (int i = 0; i < nA; i++) S[D.oA + i] = red["absa"][(size_t)i + (size_t)nA * g];
        for (int i = 0; i < nB; i++) S[D.oB + i] = red["absb"][(size_t)i + (size_t)nB * g];
        if (red.count("selfref")) for (int i = 0; i < 10; i++) S[D.oSelf + i] = red["selfref"][i + 10 * g];
        if (red.count("forref")) for (int i = 0; i < D.nfor; i++) S[D.oFor + i] = red["forref"][i + D.nfor * g];
        if (D.nsf == 1) S[D.oSflx] = red["sflux"][g];
        else for (int j = 0; j < D.nsf; j++) S[D.oSflx + j] = red["sflux"][(size_t)g + (size_t)D.ng * j];
        if (red.count("rayla")) {
          for (int j = 0; j < 9; j++) S[D.oRayl + j] = red["rayla"][(size_t)g + (size_t)D.ng * j];
          S[D.oRayl + 9] = red["raylb"][g];
        } else if (red.count("raylg")) S[D.oRayl] = red["raylg"][g];
        else S[D.oRayl] = D.rayl;
        if (red.count("absch4")) { S[D.oMisc] = red["absch4"][g]; S[D.oMisc + 1] = red["absch4"][g]; }
        if (red.count("abso3a")) { S[D.oMisc] = red["abso3a"][g]; S[D.oMisc + 1] = red["abso3b"][g]; }
        if (red.count("absco2")) { S[D.oMisc] = red["absco2"][g]; S[D.oMisc + 1] = red["absh2o"][g]; }
      }
    }
  }
  // ---------------- LW ----------------
  {
    auto R = make_reducers(16, T.get("lw_ngc"), T.get("lw_ngn"), T.get("lw_wt"));
    const auto &ngb = T.get("lw_ngb");
    for (int i = 0; i < 140; i++) T.lw_ngb[i] = (int)ngb[i];
    for (int i = 0; i < 16; i++) T.lw_delwave[i] = T.get("lw_delwave")[i];
    FileCursor F;
    if (!F.open(lw_path, err)) return ARC_ERR_IO;
    int g0 = 0;
    const char *minorA[M_COUNT] = {"ka_mn2", "ka_mn2o", "ka_mo3", "ka_mco2", "ka_mco", "ka_mo2"};
    const char *minorB[M_COUNT] = {"kb_mn2", "kb_mn2o", "kb_mo3", "kb_mco2", "kb_mco", "kb_mo2"};
    for (int ib = 0; ib < 16; ib++) {
      LwBandDesc &D = T.lw[ib];
      memset(&D, 0, sizeof(D));
      D.nspa = (int)T.get("lw_nspa")[ib]; D.nspb = (int)T.get("lw_nspb")[ib];
      D.ng = R[ib].ngc; D.g0 = g0; g0 += D.ng;
      BandRaw B;
      std::string schema = std::string(LW_SCHEMA[ib]) + LW_SF, e2;
      if (!read_band(F, schema.c_str(), D.nspa, D.nspb, B, e2)) { err = "RRTMG_LW_DATA band " + std::to_string(ib + 1) + ": " + e2; return ARC_ERR_IO; }
      std::map<std::string, std::vector<float>> red;
      std::map<std::string, int> lead;
      for (auto &kv : B.arr) {
        bool weighted = kv.first.compare(0, 7, "fracref") != 0;
        int l; R[ib].apply(kv.second, weighted, red[kv.first], l); lead[kv.first] = l;
        T.reduced["lw" + std::to_string(ib + 1) + "." + kv.first] = red[kv.first];
      }
      const int nA = lead["absa"], nB = red.count("absb") ? lead["absb"] : 0;
      D.nFracA = B.arr["fracrefa"].g_first ? lead["fracrefa"] : 1;
      D.nFracB = red.count("fracrefb") ? (B.arr["fracrefb"].g_first ? lead["fracrefb"] : 1) : 0;
      int o = 0;
      D.oA = o; o += pad4(nA + 12);
      D.oB = o; o += pad4(nB + 12);
      D.oSelf = o; o += 12;
      D.oFor = o; o += 4;
      D.oFracA = o; o += 12;
      D.oFracB = o; o += 8;
      for (int m = 0; m < M_COUNT; m++) {
        D.oMinA[m] = D.oMinB[m] = -1; D.nEtaA[m] = D.nEtaB[m] = 0;
        if (red.count(minorA[m])) { D.oMinA[m] = o; D.nEtaA[m] = lead[minorA[m]] / 19; o += pad4(lead[minorA[m]] + 1); }
        if (red.count(minorB[m])) { D.oMinB[m] = o; D.nEtaB[m] = lead[minorB[m]] / 19; o += pad4(lead[minorB[m]] + 1); }
      }
      D.oCfc = o; o += 4;
      D.slice_floats = pad4(o);
      D.slice_base = (int)T.lw_buf.size();
      T.lw_buf.resize(T.lw_buf.size() + (size_t)D.slice_floats * D.ng, 0.f);
      for (int g = 0; g < D.ng; g++) {
        float *S = &T.lw_buf[(size_t)D.slice_base + (size_t)D.slice_floats * g];
        for (int i = 0; i < nA; i++) S[D.oA + i] = red["absa"][(size_t)i + (size_t)nA * g];
        for (int i = 0; i < nB; i++) S[D.oB + i] = red["absb"][(size_t)i + (size_t)nB * g];
        for (int i = 0; i < 10; i++) S[D.oSelf + i] = red["selfref"][i + 10 * g];
        for (int i = 0; i < 4; i++) S[D.oFor + i] = red["forref"][i + 4 * g];
        if (D.nFracA == 1) S[D.oFracA] = red["fracrefa"][g];
        else for (int j = 0; j < D.nFracA; j++) S[D.oFracA + j] = red["fracrefa"][(size_t)g + (size_t)D.ng * j];
        if (D.nFracB == 1) S[D.oFracB] = red["fracrefb"][g];
        else for (int j = 0; j < D.nFracB; j++) S[D.oFracB + j] = red["fracrefb"][(size_t)g + (size_t)D.ng * j];
        for (int m = 0; m < M_COUNT; m++) {
          if (D.oMinA[m] >= 0) { int n = lead[minorA[m]]; for (int i = 0; i < n; i++) S[D.oMinA[m] + i] = red[minorA[m]][(size_t)i + (size_t)n * g]; }
          if (D.oMinB[m] >= 0) { int n = lead[minorB[m]]; for (int i = 0; i < n; i++) S[D.oMinB[m] + i] = red[minorB[m]][(size_t)i + (size_t)n * g]; }
        }
        if (red.count("ccl4")) S[D.oCfc + 0] = red["ccl4"][g];
        if (red.count("cfc11adj")) S[D.oCfc + 1] = red["cfc11adj"][g];
        if (red.count("cfc12")) S[D.oCfc + 2] = red["cfc12"][g];
        if (red.count("cfc22adj")) S[D.oCfc + 3] = red["cfc22adj"][g];
      }
    }
  }
  return 0;
}

}  // namespace arc
