// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.hpp).
// Table initialisation: restates rrtmg_sw_ini (module_ra_rrtmg_sw.F:4571-4698), the readers
// sw_kgb16..29 (SW:11315-12384), cmbgb16s..29 (SW:5022-6065), rrtmg_lw_ini
// (module_ra_rrtmg_lw.F:7862-8009), lw_kgb01..16 (LW:12956-14400) and cmbgb1..16 (LW:8206-9856).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>

#include "oracle.hpp"

namespace orc {

const FArr &InlineTables::get(const std::string &k) const {
  auto it = t.find(k);
  if (it == t.end()) {
    fprintf(stderr, "oracle: missing inline table %s\n", k.c_str());
    abort();
  }
  return it->second;
}

bool InlineTables::load(const std::string &path, std::string &err) {
  std::ifstream f(path, std::ios::binary);
  if (!f) { err = "cannot open " + path; return false; }
  std::vector<char> b((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  if (b.size() < 12 || memcmp(b.data(), "ARCTBL1\0", 8) != 0) { err = "bad magic in " + path; return false; }
  uint32_t n; memcpy(&n, &b[8], 4);
  for (uint32_t i = 0; i < n; i++) {
    const char *e = &b[12 + (size_t)i * 76];
    char name[33]; memcpy(name, e, 32); name[32] = 0;
    uint32_t nd, d[4]; int32_t lo[4]; uint64_t off;
    memcpy(&nd, e + 32, 4); memcpy(d, e + 36, 16); memcpy(lo, e + 52, 16); memcpy(&off, e + 68, 8);
    FArr a; a.alloc(d[0], d[1], d[2], d[3]);
    for (int q = 0; q < 4; q++) a.lo[q] = lo[q];
    if (off + a.size() * 4 > b.size()) { err = "truncated table file"; return false; }
    memcpy(a.v.data(), &b[off], a.size() * 4);
    t[name] = a;
  }
  return true;
}

Tables &tables() { static Tables T; return T; }

namespace {

struct RecReader {
  std::vector<char> buf; size_t pos = 0; size_t rec_end = 0; std::string err;
  bool open(const std::string &p) {
    std::ifstream f(p, std::ios::binary);
    if (!f) { err = "cannot open " + p; return false; }
    buf.assign((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    return true;
  }
  bool begin() {
    if (pos + 4 > buf.size()) { err = "unexpected end of k-table file"; return false; }
    int32_t n; memcpy(&n, &buf[pos], 4); pos += 4; rec_end = pos + n;
    if (rec_end + 4 > buf.size()) { err = "truncated record"; return false; }
    return true;
  }
  bool end() {
    if (pos != rec_end) { err = "record length mismatch (" + std::to_string((long)rec_end - (long)pos) + " bytes left)"; return false; }
    int32_t n; memcpy(&n, &buf[pos], 4); pos += 4; return true;
  }
  float f32() { float x; memcpy(&x, &buf[pos], 4); pos += 4; return x; }
  int i32() { int32_t x; memcpy(&x, &buf[pos], 4); pos += 4; return x; }
  FArr arr(int a, int b = 1, int c = 1, int d = 1) {
    FArr r; r.alloc(a, b, c, d);
    if (pos + r.size() * 4 > rec_end) { err = "record too short"; r.v.assign(r.size(), 0.f); pos = rec_end + 1; return r; }
    memcpy(r.v.data(), &buf[pos], r.size() * 4); pos += r.size() * 4; return r;
  }
};

// g-point reduction of one array of band `ibnd` (1-based in the SW 1..14 / LW 1..16 numbering):
// the loop nest of every cmbgbNN routine (e.g. SW:5037-5051): for each reduced g-point igc sum
// the ngn(igc) original points, weighted by rwgt(iprsm+16*(ibnd-1)) unless the array is a
// source term (sfluxrefo / fracrefao / fracrefbo are summed unweighted, SW:5103-5111, LW:8302-8313).
FArr reduce_g(const FArr &raw, int ibnd, const int *ngc, const int *ngs, const float *ngn, const float *rwgt,
              bool weighted, bool g_first) {
  int ng_out = ngc[ibnd - 1];
  int goff = ibnd >= 2 ? ngs[ibnd - 2] : 0;
  FArr out;
  if (g_first) {
    int ncol = (int)(raw.size() / 16);
    out.alloc(ng_out, ncol);
    for (int j = 0; j < ncol; j++) {
      int iprsm = 0;
      for (int igc = 1; igc <= ng_out; igc++) {
        float s = 0.f;
        int cnt = (int)ngn[goff + igc - 1];
        for (int ipr = 1; ipr <= cnt; ipr++) {
          iprsm++;
          float x = raw.v[(iprsm - 1) + 16 * (size_t)j];
          s = s + (weighted ? x * rwgt[iprsm - 1 + 16 * (ibnd - 1)] : x);
        }
        out.v[(igc - 1) + (size_t)ng_out * j] = s;
      }
    }
  } else {
    size_t lead = raw.size() / 16;
    out.alloc((int)lead, ng_out);
    for (size_t l = 0; l < lead; l++) {
      int iprsm = 0;
      for (int igc = 1; igc <= ng_out; igc++) {
        float s = 0.f;
        int cnt = (int)ngn[goff + igc - 1];
        for (int ipr = 1; ipr <= cnt; ipr++) {
          iprsm++;
          float x = raw.v[l + lead * (size_t)(iprsm - 1)];
          s = s + (weighted ? x * rwgt[iprsm - 1 + 16 * (ibnd - 1)] : x);
        }
        out.v[l + lead * (size_t)(igc - 1)] = s;
      }
    }
  }
  return out;
}

// shape helper: reduced array as 2-D (lead, ng) / 3-D (a, b, ng)
void shape2(FArr &a, int n0, int n1) { a.n[0] = n0; a.n[1] = n1; a.n[2] = 1; a.n[3] = 1; }
void shape3(FArr &a, int n0, int n1, int n2) { a.n[0] = n0; a.n[1] = n1; a.n[2] = n2; a.n[3] = 1; }

// relative weights, rrtmg_sw_ini SW:4656-4680 / rrtmg_lw_ini LW:7973-7997
void make_rwgt(int nb, const int *ngc, const float *ngn, const float *ngm, const float *wt, float *rwgt) {
  int igcsm = 0;
  float wtsm[17];
  for (int ibnd = 1; ibnd <= nb; ibnd++) {
    int iprsm = 0;
    if (ngc[ibnd - 1] < 16) {
      for (int igc = 1; igc <= ngc[ibnd - 1]; igc++) {
        igcsm++;
        float wtsum = 0.f;
        for (int ipr = 1; ipr <= (int)ngn[igcsm - 1]; ipr++) { iprsm++; wtsum = wtsum + wt[iprsm - 1]; }
        wtsm[igc] = wtsum;
      }
      for (int ig = 1; ig <= 16; ig++) {
        int ind = (ibnd - 1) * 16 + ig;
        rwgt[ind - 1] = wt[ig - 1] / wtsm[(int)ngm[ind - 1]];
      }
    } else {
      for (int ig = 1; ig <= 16; ig++) {
        igcsm++;
        int ind = (ibnd - 1) * 16 + ig;
        rwgt[ind - 1] = 1.0f;
      }
    }
  }
}

}  // namespace

int init_tables(const std::string &inline_path, const std::string &sw_path, const std::string &lw_path,
                float cp, float p_top, int kme, std::string &err) {
  Tables &T = tables();
  T = Tables();
  if (!T.in.load(inline_path, err)) return 3;
  // ---- constants: swdatinit SW:4763-4788, lwdatinit LW:8096-8120
  T.oneminus = 1.0f - 1.e-06f;
  T.pi = 2.0f * asinf(1.0f);
  T.heatfac = T.grav * T.secdy / (cp * 1.e2f);
  T.fluxfac = T.pi * 2.e4f;
  // LW:12861  NLAYERS = kme + nint(p_top*0.01/deltap) - 1   (Fortran NINT: half away from zero)
  T.lw_nlayers = kme + (int)lroundf(p_top * 0.01f / 4.0f) - 1;

  // ================= SW =================
  const FArr &ngc = T.in.get("sw_ngc"), &ngs = T.in.get("sw_ngs"), &ngn = T.in.get("sw_ngn"),
             &ngm = T.in.get("sw_ngm"), &ngb = T.in.get("sw_ngb"), &wt = T.in.get("sw_wt");
  for (int i = 0; i < 14; i++) {
    T.sw_ngc[i] = (int)ngc.v[i]; T.sw_ngs[i] = (int)ngs.v[i];
    T.sw_nspa[i] = (int)T.in.get("sw_nspa").v[i]; T.sw_nspb[i] = (int)T.in.get("sw_nspb").v[i];
  }
  for (int i = 0; i < 112; i++) T.sw_ngb[i] = (int)ngb.v[i];
  make_rwgt(14, T.sw_ngc, ngn.v.data(), ngm.v.data(), wt.v.data(), T.sw_rwgt);
  // exp_tbl SW:4640-4649
  {
    const int ntbl = 10000; const float pade = 0.278f, expeps = 1.e-20f;
    T.sw_exp_tbl.assign(ntbl + 1, 0.f);
    T.sw_exp_tbl[0] = 1.0f; T.sw_exp_tbl[ntbl] = expeps;
    T.sw_bpade = 1.0f / pade;
    for (int itr = 1; itr <= ntbl - 1; itr++) {
      float tfn = (float)itr / (float)ntbl;
      float tau_tbl = T.sw_bpade * tfn / (1.0f - tfn);
      T.sw_exp_tbl[itr] = expf(-tau_tbl);
      if (T.sw_exp_tbl[itr] <= expeps) T.sw_exp_tbl[itr] = expeps;
    }
  }
  {
    RecReader R;
    if (!R.open(sw_path)) { err = R.err; return 3; }
    auto red = [&](const FArr &raw, int ib, bool w = true, bool gf = false) {
      return reduce_g(raw, ib, T.sw_ngc, T.sw_ngs, ngn.v.data(), T.sw_rwgt, w, gf);
    };
    for (int ib = 1; ib <= 14; ib++) {
      SwBand &B = T.sw[ib - 1];
      int band = ib + 15;
      B.nspa = T.sw_nspa[ib - 1]; B.nspb = T.sw_nspb[ib - 1]; B.ng = T.sw_ngc[ib - 1];
      if (!R.begin()) { err = R.err; return 3; }
      FArr kao, kbo, selfo, foro, sflo, raylo, raylao, raylbo, o3a, o3b, ch4o, h2oo, co2o;
      auto KA = [&]() { kao = B.nspa > 1 ? R.arr(B.nspa, 5, 13, 16) : R.arr(5, 13, 16); };
      auto KB = [&]() { kbo = B.nspb > 1 ? R.arr(B.nspb, 5, 47, 16) : R.arr(5, 47, 16); };
      switch (band) {
        case 16: B.rayl = R.f32(); B.strrat = R.f32(); B.layreffr = R.i32(); KA(); KB();
                 selfo = R.arr(10, 16); foro = R.arr(3, 16); sflo = R.arr(16); B.nsf = 1; break;
        case 17: B.rayl = R.f32(); B.strrat = R.f32(); B.layreffr = R.i32(); KA(); KB();
                 selfo = R.arr(10, 16); foro = R.arr(4, 16); sflo = R.arr(16, 5); B.nsf = 5; break;
        case 18: case 19: case 22:
                 B.rayl = R.f32(); B.strrat = R.f32(); B.layreffr = R.i32(); KA(); KB();
                 selfo = R.arr(10, 16); foro = R.arr(3, 16); sflo = R.arr(16, 9); B.nsf = 9; break;
        case 20: B.rayl = R.f32(); B.layreffr = R.i32(); ch4o = R.arr(16); KA(); KB();
                 selfo = R.arr(10, 16); foro = R.arr(4, 16); sflo = R.arr(16); B.nsf = 1; break;
        case 21: B.rayl = R.f32(); B.strrat = R.f32(); B.layreffr = R.i32(); KA(); KB();
                 selfo = R.arr(10, 16); foro = R.arr(4, 16); sflo = R.arr(16, 9); B.nsf = 9; break;
        case 23: raylo = R.arr(16); B.givfac = R.f32(); B.layreffr = R.i32(); KA();
                 selfo = R.arr(10, 16); foro = R.arr(3, 16); sflo = R.arr(16); B.nsf = 1; break;
        case 24: raylao = R.arr(16, 9); raylbo = R.arr(16); B.strrat = R.f32(); B.layreffr = R.i32();
                 o3a = R.arr(16); o3b = R.arr(16); KA(); KB();
                 selfo = R.arr(10, 16); foro = R.arr(3, 16); sflo = R.arr(16, 9); B.nsf = 9; break;
        case 25: raylo = R.arr(16); B.layreffr = R.i32(); o3a = R.arr(16); o3b = R.arr(16); KA();
                 sflo = R.arr(16); B.nsf = 1; break;
        case 26: raylo = R.arr(16); sflo = R.arr(16); B.nsf = 1; break;
        case 27: raylo = R.arr(16); B.scalekur = R.f32(); B.layreffr = R.i32(); KA(); KB();
                 sflo = R.arr(16); B.nsf = 1; break;
        case 28: B.rayl = R.f32(); B.strrat = R.f32(); B.layreffr = R.i32(); KA(); KB();
                 sflo = R.arr(16, 5); B.nsf = 5; break;
        case 29: B.rayl = R.f32(); B.layreffr = R.i32(); h2oo = R.arr(16); co2o = R.arr(16); KA(); KB();
                 selfo = R.arr(10, 16); foro = R.arr(4, 16); sflo = R.arr(16); B.nsf = 1; break;
      }
      if (!R.err.empty() || !R.end()) { err = "RRTMG_SW_DATA band " + std::to_string(band) + ": " + R.err; return 3; }
      if (kao.size()) { B.absa = red(kao, ib); shape2(B.absa, 65 * std::max(B.nspa, 1), B.ng); }
      if (kbo.size()) { B.absb = red(kbo, ib); shape2(B.absb, 235 * std::max(B.nspb, 1), B.ng); }
      if (selfo.size()) { B.selfref = red(selfo, ib); shape2(B.selfref, 10, B.ng); }
      if (foro.size()) { B.nfor = foro.n[0]; B.forref = red(foro, ib); shape2(B.forref, B.nfor, B.ng); }
      if (B.nsf == 1) { B.sfluxref = red(sflo, ib, false); shape2(B.sfluxref, B.ng, 1); }
      else { B.sfluxref = red(sflo, ib, false, true); }
      if (raylo.size()) { B.raylg = red(raylo, ib); shape2(B.raylg, B.ng, 1); }
      if (raylao.size()) { B.rayla = red(raylao, ib, true, true); }
      if (raylbo.size()) { B.raylb = red(raylbo, ib); shape2(B.raylb, B.ng, 1); }
      if (o3a.size()) { B.abso3a = red(o3a, ib); shape2(B.abso3a, B.ng, 1); B.abso3b = red(o3b, ib); shape2(B.abso3b, B.ng, 1); }
      if (ch4o.size()) { B.absch4 = red(ch4o, ib); shape2(B.absch4, B.ng, 1); }
      if (h2oo.size()) { B.absh2o = red(h2oo, ib); shape2(B.absh2o, B.ng, 1); B.absco2 = red(co2o, ib); shape2(B.absco2, B.ng, 1); }
    }
  }

  // ================= LW =================
  const FArr &lngc = T.in.get("lw_ngc"), &lngs = T.in.get("lw_ngs"), &lngn = T.in.get("lw_ngn"),
             &lngm = T.in.get("lw_ngm"), &lngb = T.in.get("lw_ngb"), &lwt = T.in.get("lw_wt");
  for (int i = 0; i < 16; i++) {
    T.lw_ngc[i] = (int)lngc.v[i]; T.lw_ngs[i] = (int)lngs.v[i];
    T.lw_nspa[i] = (int)T.in.get("lw_nspa").v[i]; T.lw_nspb[i] = (int)T.in.get("lw_nspb").v[i];
    T.lw_delwave[i] = T.in.get("lw_delwave").v[i];
  }
  for (int i = 0; i < 140; i++) T.lw_ngb[i] = (int)lngb.v[i];
  make_rwgt(16, T.lw_ngc, lngn.v.data(), lngm.v.data(), lwt.v.data(), T.lw_rwgt);
  // tau_tbl / exp_tbl / tfn_tbl  LW:7940-7957
  {
    const int ntbl = 10000; const float pade = 0.278f, expeps = 1.e-20f;
    T.lw_tau_tbl.assign(ntbl + 1, 0.f); T.lw_exp_tbl.assign(ntbl + 1, 0.f); T.lw_tfn_tbl.assign(ntbl + 1, 0.f);
    T.lw_tau_tbl[0] = 0.0f; T.lw_tau_tbl[ntbl] = 1.e10f;
    T.lw_exp_tbl[0] = 1.0f; T.lw_exp_tbl[ntbl] = expeps;
    T.lw_tfn_tbl[0] = 0.0f; T.lw_tfn_tbl[ntbl] = 1.0f;
    T.lw_bpade = 1.0f / pade;
    for (int itr = 1; itr <= ntbl - 1; itr++) {
      float tfn = (float)itr / (float)ntbl;
      T.lw_tau_tbl[itr] = T.lw_bpade * tfn / (1.0f - tfn);
      T.lw_exp_tbl[itr] = expf(-T.lw_tau_tbl[itr]);
      if (T.lw_exp_tbl[itr] <= expeps) T.lw_exp_tbl[itr] = expeps;
      if (T.lw_tau_tbl[itr] < 0.06f) T.lw_tfn_tbl[itr] = T.lw_tau_tbl[itr] / 6.0f;
      else T.lw_tfn_tbl[itr] = 1.0f - 2.0f * ((1.0f / T.lw_tau_tbl[itr]) - (T.lw_exp_tbl[itr] / (1.0f - T.lw_exp_tbl[itr])));
    }
  }
  {
    RecReader R;
    if (!R.open(lw_path)) { err = R.err; return 3; }
    auto red = [&](const FArr &raw, int ib, bool w = true, bool gf = false) {
      return reduce_g(raw, ib, T.lw_ngc, T.lw_ngs, lngn.v.data(), T.lw_rwgt, w, gf);
    };
    for (int ib = 1; ib <= 16; ib++) {
      LwBand &B = T.lw[ib - 1];
      B.nspa = T.lw_nspa[ib - 1]; B.nspb = T.lw_nspb[ib - 1]; B.ng = T.lw_ngc[ib - 1];
      const int ng = B.ng;
      if (!R.begin()) { err = R.err; return 3; }
      FArr fa, fb, kao, kbo, selfo, foro;
      auto KA = [&]() { kao = B.nspa > 1 ? R.arr(B.nspa, 5, 13, 16) : R.arr(5, 13, 16); };
      auto KB = [&]() { kbo = B.nspb > 1 ? R.arr(B.nspb, 5, 47, 16) : R.arr(5, 47, 16); };
      auto SF = [&]() { selfo = R.arr(10, 16); foro = R.arr(4, 16); };
      auto m1 = [&](FArr &dst) { FArr r = R.arr(19, 16); dst = red(r, ib); shape2(dst, 19, ng); };
      auto m2 = [&](FArr &dst, int ne) { FArr r = R.arr(ne, 19, 16); dst = red(r, ib); shape3(dst, ne, 19, ng); };
      auto v1 = [&](FArr &dst) { FArr r = R.arr(16); dst = red(r, ib); shape2(dst, ng, 1); };
      bool fa9 = false, fb5 = false, hasfb = true;
      switch (ib) {
        case 1: fa = R.arr(16); fb = R.arr(16); KA(); KB(); m1(B.ka_mn2); m1(B.kb_mn2); SF(); break;
        case 2: case 10: case 14: fa = R.arr(16); fb = R.arr(16); KA(); KB(); SF(); break;
        case 3: fa = R.arr(16, 9); fb = R.arr(16, 5); fa9 = fb5 = true; KA(); KB(); m2(B.ka_mn2o, 9); m2(B.kb_mn2o, 5); SF(); break;
        case 4: fa = R.arr(16, 9); fb = R.arr(16, 5); fa9 = fb5 = true; KA(); KB(); SF(); break;
        case 5: fa = R.arr(16, 9); fb = R.arr(16, 5); fa9 = fb5 = true; KA(); KB(); m2(B.ka_mo3, 9); v1(B.ccl4); SF(); break;
        case 6: fa = R.arr(16); hasfb = false; KA(); m1(B.ka_mco2); v1(B.cfc11adj); v1(B.cfc12); SF(); break;
        case 7: fa = R.arr(16, 9); fa9 = true; fb = R.arr(16); KA(); KB(); m2(B.ka_mco2, 9); m1(B.kb_mco2); SF(); break;
        case 8: fa = R.arr(16); fb = R.arr(16); KA(); KB(); m1(B.ka_mco2); m1(B.kb_mco2); m1(B.ka_mn2o); m1(B.kb_mn2o);
                m1(B.ka_mo3); v1(B.cfc12); v1(B.cfc22adj); SF(); break;
        case 9: fa = R.arr(16, 9); fa9 = true; fb = R.arr(16); KA(); KB(); m2(B.ka_mn2o, 9); m1(B.kb_mn2o); SF(); break;
        case 11: fa = R.arr(16); fb = R.arr(16); KA(); KB(); m1(B.ka_mo2); m1(B.kb_mo2); SF(); break;
        case 12: fa = R.arr(16, 9); fa9 = true; hasfb = false; KA(); SF(); break;
        case 13: fa = R.arr(16, 9); fa9 = true; fb = R.arr(16); KA(); m2(B.ka_mco2, 9); m2(B.ka_mco, 9); m1(B.kb_mo3); SF(); break;
        case 15: fa = R.arr(16, 9); fa9 = true; hasfb = false; KA(); m2(B.ka_mn2, 9); SF(); break;
        case 16: fa = R.arr(16, 9); fa9 = true; fb = R.arr(16); KA(); KB(); SF(); break;
      }
      if (!R.err.empty() || !R.end()) { err = "RRTMG_LW_DATA band " + std::to_string(ib) + ": " + R.err; return 3; }
      B.absa = red(kao, ib); shape2(B.absa, 65 * B.nspa, ng);
      if (kbo.size()) { B.absb = red(kbo, ib); shape2(B.absb, 235 * std::max(B.nspb, 1), ng); }
      B.selfref = red(selfo, ib); shape2(B.selfref, 10, ng);
      B.forref = red(foro, ib); shape2(B.forref, 4, ng);
      if (fa9) B.fracrefa = red(fa, ib, false, true); else { B.fracrefa = red(fa, ib, false); shape2(B.fracrefa, ng, 1); }
      if (hasfb) { if (fb5) B.fracrefb = red(fb, ib, false, true); else { B.fracrefb = red(fb, ib, false); shape2(B.fracrefb, ng, 1); } }
    }
  }
  T.ready = true;
  return 0;
}

}  // namespace orc
