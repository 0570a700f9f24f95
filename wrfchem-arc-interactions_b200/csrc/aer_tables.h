// Chebyshev-Mie coefficient tables of the aerosol optical-property stage (see aer_tables.cpp).
#pragma once
#include <vector>

namespace arc {

constexpr int AER_NWL = 20;         // 4 SW wavelengths (300, 400, 600, 999 nm) + 16 RRTMG-LW band centres
constexpr int AER_NSW = 4, AER_NLW = 16;
constexpr int AER_NCLASS = 9;       // so4 no3 cl nh4 na oin oc bc water
constexpr int AER_NQ = 3;           // ln Q_ext, ln Q_sca, ln g
constexpr int AER_NREFR = 7, AER_NREFI = 7, AER_NCOEF = 50, AER_NCOEF_PAD = 52, AER_NSIZ = 200;
constexpr int AER_MAXBIN = 8, AER_MAXSPEC = 24;

struct AerTables {
  double wavelength_cm[AER_NWL];
  float nr[AER_NCLASS][AER_NWL], ni[AER_NCLASS][AER_NWL];      // species refractive indices n + i k
  float refr_lo[AER_NWL], refr_hi[AER_NWL], refi_lo[AER_NWL], refi_hi[AER_NWL];
  double rmin, rmax;                                           // cm
  std::vector<float> coef;                                     // [wl][q][refr][refi][AER_NCOEF_PAD]
};

void default_refindex(float nr[AER_NCLASS][AER_NWL], float ni[AER_NCLASS][AER_NWL]);
int build_aer_tables(const float nr[AER_NCLASS][AER_NWL], const float ni[AER_NCLASS][AER_NWL], AerTables &T);
void mie_efficiencies(double x, double nr, double ni, double &qext, double &qsca, double &asym);

}  // namespace arc
