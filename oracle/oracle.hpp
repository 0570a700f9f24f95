// ORACLE -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement (C++17, FP32, one column per call like the reference's ncol=1) of the
// v3.9.1 RRTMG SW/LW path of douglowe/WRFChem-ARC-Interactions.  Nothing under oracle/ is
// part of the shipped product: only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it, and only as the checker / baseline.
//
// PARITY PIN STATUS: the reference holds no golden vectors and no Fortran compiler exists
// in the build container, so this restatement is pinned by (a) the reference's six
// structural invariants and (b) line-by-line review against the
// cited Fortran.  Absolute fluxes are "parity unpinned" with respect to a gfortran run.
//
// Build flags matter: -O2 -ffp-contract=off -fno-fast-math (no FMA contraction), so index
// arithmetic is what gfortran -O3 produces on baseline x86-64.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace orc {

// Fortran-ordered (column-major, 1-based by default) float array with up to 4 dims.
struct FArr {
  std::vector<float> v;
  int n[4] = {1, 1, 1, 1};
  int lo[4] = {1, 1, 1, 1};
  void alloc(int a, int b = 1, int c = 1, int d = 1) {
    n[0] = a; n[1] = b; n[2] = c; n[3] = d;
    v.assign((size_t)a * b * c * d, 0.f);
  }
  inline float &operator()(int i) { return v[i - lo[0]]; }
  inline float operator()(int i) const { return v[i - lo[0]]; }
  inline float &operator()(int i, int j) { return v[(i - lo[0]) + (size_t)n[0] * (j - lo[1])]; }
  inline float operator()(int i, int j) const { return v[(i - lo[0]) + (size_t)n[0] * (j - lo[1])]; }
  inline float &operator()(int i, int j, int k) {
    return v[(i - lo[0]) + (size_t)n[0] * ((j - lo[1]) + (size_t)n[1] * (k - lo[2]))];
  }
  inline float operator()(int i, int j, int k) const {
    return v[(i - lo[0]) + (size_t)n[0] * ((j - lo[1]) + (size_t)n[1] * (k - lo[2]))];
  }
  size_t size() const { return v.size(); }
};

// Inline (source-embedded) tables of the reference, loaded from rrtmg_inline_tables.bin
struct InlineTables {
  std::map<std::string, FArr> t;
  const FArr &get(const std::string &k) const;
  bool load(const std::string &path, std::string &err);
};

// One spectral band after g-point reduction (reference modules rrsw_kgNN / rrlw_kgNN)
struct SwBand {
  int nspa = 0, nspb = 0, ng = 0, nfor = 0, nsf = 1, layreffr = 0;
  float rayl = 0, strrat = 0, givfac = 0, scalekur = 0;
  FArr absa, absb, selfref, forref, sfluxref;  // absa(65*nspa, ng) absb(235*nspb, ng) selfref(10,ng) forref(nfor,ng) sfluxref(ng,nsf)
  FArr raylg, rayla, raylb, abso3a, abso3b, absch4, absh2o, absco2;
};

struct LwBand {
  int nspa = 0, nspb = 0, ng = 0;
  FArr absa, absb, selfref, forref;      // forref(4,ng)
  FArr fracrefa, fracrefb;               // (ng) or (ng,9)/(ng,5)
  FArr ka_mn2, kb_mn2;                   // band 1: (19,ng); band 15 ka_mn2(9,19,ng)
  FArr ka_mn2o, kb_mn2o;                 // band 3: (9,19,ng),(5,19,ng); 8: (19,ng) both; 9: (9,19,ng),(19,ng)
  FArr ka_mo3, kb_mo3;                   // band 5 ka(9,19,ng); 8 ka(19,ng); 13 kb(19,ng)
  FArr ka_mco2, kb_mco2;                 // band 6 ka(19,ng); 7 ka(9,19,ng), kb(19,ng); 8 both (19,ng); 13 ka(9,19,ng)
  FArr ka_mco;                           // band 13 (9,19,ng)
  FArr ka_mo2, kb_mo2;                   // band 11 (19,ng)
  FArr ccl4, cfc11adj, cfc12, cfc22adj;  // (ng)
};

struct Tables {
  InlineTables in;
  // constants (swdatinit SW:4701-4790, lwdatinit LW:8012-8122)
  float grav = 9.8066f, avogad = 6.02214199e+23f, secdy = 8.6400e4f;
  float oneminus, pi, heatfac, fluxfac;
  // SW
  int sw_ngc[14], sw_ngs[14], sw_ngb[112], sw_nspa[14], sw_nspb[14];
  float sw_rwgt[224];
  SwBand sw[14];
  std::vector<float> sw_exp_tbl;  // 0..10000
  float sw_bpade;
  // LW
  int lw_ngc[16], lw_ngs[16], lw_ngb[140], lw_nspa[16], lw_nspb[16];
  float lw_rwgt[256], lw_delwave[16];
  LwBand lw[16];
  std::vector<float> lw_tau_tbl, lw_exp_tbl, lw_tfn_tbl;
  float lw_bpade;
  int lw_nlayers = 0;  // module variable set by rrtmg_lwinit (LW:12861)
  bool ready = false;
};

Tables &tables();
int init_tables(const std::string &inline_path, const std::string &sw_path, const std::string &lw_path,
                float cp, float p_top, int kme, std::string &err);

static const int MXLAY = 260;
static const int NGSW = 112, NBSW = 14, NGLW = 140, NBLW = 16;

}  // namespace orc
