// Host-side table preparation for the B200 radiation path (init, row a18 of SURVEY.md section 8).
//
// What the reference does at init (rrtmg_sw_ini module_ra_rrtmg_sw.F:4571-4698, sw_kgb16..29
// SW:11315-12384, cmbgb16s..29 SW:5022-6065; rrtmg_lw_ini module_ra_rrtmg_lw.F:7862-8009,
// lw_kgb01..16 LW:12956-14400, cmbgb1..16 LW:8206-9856) is done here once on the host:
//   1. parse the Fortran sequential-unformatted records of RRTMG_SW_DATA / RRTMG_LW_DATA,
//   2. reduce every 16-g array to the band's ngc g-points (weights rwgt = wt/sum(wt); source
//      terms sfluxrefo / fracref*o are plain sums),
//   3. re-lay the reduced tables out as one contiguous "slice" per (band, g-point) so a g-point's
//      whole working set is a single 16-byte-aligned TMA bulk copy (cp.async.bulk) into shared memory.
#pragma once
#include <map>
#include <string>
#include <vector>

namespace arc {

// offsets (in floats) inside a slice; -1 = absent
struct SwBandDesc {
  int ng, g0, nspa, nspb, layreffr, nsf, nfor;
  int slice_floats;   // stride between consecutive g slices (multiple of 4 floats)
  int slice_base;     // float offset of this band's g=0 slice inside the SW table buffer
  int oA, oB, oSelf, oFor, oSflx, oRayl, oMisc;
  float rayl, strrat, givfac, scalekur;
};

// LW minor-species table slots inside a slice
enum LwMinor { M_N2 = 0, M_N2O, M_O3, M_CO2, M_CO, M_O2, M_COUNT };

struct LwBandDesc {
  int ng, g0, nspa, nspb;
  int slice_floats, slice_base;
  int oA, oB, oSelf, oFor, oFracA, oFracB;   // fracA: 9 entries (eta) or 1; fracB: 5 or 1
  int nFracA, nFracB;
  int oMinA[M_COUNT], oMinB[M_COUNT];        // lower / upper minor tables, -1 absent
  int nEtaA[M_COUNT], nEtaB[M_COUNT];        // leading eta dimension (1 = T-only table of 19)
  int oCfc;                                  // 4 floats: ccl4, cfc11adj, cfc12, cfc22adj
};

struct HostTables {
  // inline tables (rrtmg_inline_tables.bin)
  std::map<std::string, std::vector<float>> in;
  std::map<std::string, std::vector<int>> in_dims;
  const std::vector<float> &get(const std::string &k) const;

  float cp = 1004.5f, heatfac = 0, fluxfac = 0, oneminus = 0, pi = 0;
  int lw_nlayers = 0;

  SwBandDesc sw[14];
  std::vector<float> sw_buf;      // all SW slices
  int sw_ngb[112];                // band (1..14) of each g-point
  LwBandDesc lw[16];
  std::vector<float> lw_buf;
  int lw_ngb[140];
  float lw_delwave[16];

  std::vector<float> sw_exp_tbl;                     // 10001
  std::vector<float> lw_tau_tbl, lw_exp_tbl, lw_tfn_tbl;
  float bpade = 0;

  // reduced tables by name ("sw16.absa" ...), kept for the table-parity taps
  std::map<std::string, std::vector<float>> reduced;
};

// returns 0 or an ARC_ERR_* code; fills err
int build_host_tables(const std::string &inline_path, const std::string &sw_path, const std::string &lw_path,
                      float cp, float p_top, int kme, HostTables &T, std::string &err);

}  // namespace arc
