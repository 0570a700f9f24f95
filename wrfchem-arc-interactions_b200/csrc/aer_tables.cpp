// Host-side construction of the Chebyshev-Mie coefficient tables of the aerosol optical-property stage.
//
// Upstream algorithm (WRF-Chem v3.9.1 chem/module_optical_averaging.F, subroutine mieaer; NOT part of the reference
// repository -- SURVEY.md section 0.4 -- restated from its published description: Fast et al. 2006, Barnard et al. 2010,
// Ghan & Zaveri 2007): for every wavelength and every node of a 7 x 7 (n_r, n_i) refractive-index grid, full Mie theory
// is evaluated at 200 Chebyshev nodes in ln(r) between 0.005 and 50 um, and ln(Q_ext), ln(Q_sca), ln(g) are each fitted
// with a 50-term Chebyshev series; at run time the coefficients are interpolated bilinearly in (n_r, n_i) and the series
// evaluated at the section's wet radius.
#include "aer_tables.h"

#include <algorithm>
#include <cmath>
#include <complex>

namespace arc {

namespace {
typedef std::complex<double> cd;

// Mie efficiencies of a homogeneous sphere: size parameter x, refractive index n + i k (k >= 0 absorbing).
// Logarithmic derivative D_n(mx) by downward recurrence, Riccati-Bessel psi/xi by upward recurrence (Bohren & Huffman 1983).
void mie_sphere(double x, double nr, double ni, double &qext, double &qsca, double &asym) {
  const cd m(nr, ni);                 // Bohren-Huffman sign convention: m = n + i k with xi = psi - i chi
  const cd y = m * x;
  const int nstop = (int)(x + 4.0 * std::cbrt(x) + 2.0);
  const int nmx = (int)std::max((double)nstop, std::abs(y)) + 15;
  std::vector<cd> D(nmx + 1);
  D[nmx] = cd(0.0, 0.0);
  for (int n = nmx; n >= 1; n--) {
    const cd t = (double)n / y;
    D[n - 1] = t - 1.0 / (D[n] + t);
  }
  double psi0 = std::cos(x), psi1 = std::sin(x), chi0 = -std::sin(x), chi1 = std::cos(x);
  cd xi1(psi1, -chi1);
  double qs = 0.0, gs = 0.0, qe = 0.0;
  cd an1(0, 0), bn1(0, 0);
  for (int n = 1; n <= nstop; n++) {
    const double fn = (2.0 * n + 1.0) / (n * (n + 1.0));
    const double psi = (2.0 * n - 1.0) * psi1 / x - psi0;
    const double chi = (2.0 * n - 1.0) * chi1 / x - chi0;
    const cd xi(psi, -chi);
    const cd da = D[n] / m + (double)n / x, db = m * D[n] + (double)n / x;
    const cd an = (da * psi - psi1) / (da * xi - xi1);
    const cd bn = (db * psi - psi1) / (db * xi - xi1);
    qs += (2.0 * n + 1.0) * (std::norm(an) + std::norm(bn));
    qe += (2.0 * n + 1.0) * (an.real() + bn.real());
    gs += fn * (an * std::conj(bn)).real();
    if (n > 1) gs += ((n - 1.0) * (n + 1.0) / n) * ((an1 * std::conj(an)).real() + (bn1 * std::conj(bn)).real());
    psi0 = psi1; psi1 = psi; chi0 = chi1; chi1 = chi; xi1 = cd(psi1, -chi1);
    an1 = an; bn1 = bn;
  }
  qsca = 2.0 / (x * x) * qs;
  qext = 2.0 / (x * x) * qe;
  asym = qs > 0.0 ? 2.0 * gs / qs : 0.0;      // g = (4/x^2) gs / Q_sca
}
}  // namespace

void mie_efficiencies(double x, double nr, double ni, double &qext, double &qsca, double &asym) { mie_sphere(x, nr, ni, qext, qsca, asym); }

// default complex refractive indices by species class and wavelength.  Real data live in WRF-Chem's
// module_data_rrtmgaeropt.F, which is not in the reference repository: the values below are representative literature
// numbers (sulfate-like, sea-salt-like, mineral dust, organic carbon, soot, water) and can be replaced through ArcAerConfig.
void default_refindex(float nr[AER_NCLASS][AER_NWL], float ni[AER_NCLASS][AER_NWL]) {
  // SW wavelengths 0.30 0.40 0.60 0.999 um, then the 16 RRTMG-LW band centres (um) 55.6 23.5 17.7 15.0 13.2 11.1 9.71 8.85 7.78 6.97 6.10 5.15 4.62 4.32 4.02 3.42
  const float w_nr[AER_NWL] = {1.349f, 1.339f, 1.332f, 1.327f, 1.55f, 1.50f, 1.42f, 1.27f, 1.15f, 1.16f, 1.22f, 1.26f, 1.30f, 1.32f, 1.33f, 1.32f, 1.33f, 1.33f, 1.35f, 1.42f};
  const float w_ni[AER_NWL] = {1.6e-8f, 1.9e-9f, 1.1e-8f, 2.9e-6f, 0.50f, 0.39f, 0.43f, 0.40f, 0.30f, 0.10f, 0.05f, 0.04f, 0.035f, 0.032f, 0.11f, 0.012f, 0.015f, 0.009f, 0.005f, 0.02f};
  const float s_nr[AER_NWL] = {1.47f, 1.44f, 1.43f, 1.42f, 1.89f, 1.91f, 1.93f, 1.59f, 1.59f, 1.72f, 1.89f, 1.67f, 1.22f, 1.36f, 1.42f, 1.34f, 1.34f, 1.34f, 1.37f, 1.40f};
  const float s_ni[AER_NWL] = {1e-8f, 1e-8f, 1e-8f, 1.7e-6f, 0.22f, 0.15f, 0.26f, 0.32f, 0.28f, 0.31f, 0.46f, 0.63f, 0.15f, 0.10f, 0.09f, 0.13f, 0.12f, 0.12f, 0.13f, 0.16f};
  const float d_nr[AER_NWL] = {1.55f, 1.55f, 1.55f, 1.55f, 2.34f, 2.90f, 1.75f, 1.51f, 1.62f, 1.82f, 2.92f, 1.35f, 1.19f, 1.42f, 1.43f, 1.45f, 1.46f, 1.46f, 1.47f, 1.48f};
  const float d_ni[AER_NWL] = {0.003f, 0.003f, 0.003f, 0.003f, 0.70f, 0.86f, 0.41f, 0.22f, 0.20f, 0.34f, 0.65f, 0.20f, 0.10f, 0.06f, 0.06f, 0.01f, 0.007f, 0.006f, 0.005f, 0.004f};
  const float n_nr[AER_NWL] = {1.51f, 1.50f, 1.49f, 1.47f, 1.74f, 1.76f, 1.76f, 1.62f, 1.51f, 1.48f, 1.56f, 1.60f, 1.40f, 1.42f, 1.45f, 1.47f, 1.47f, 1.47f, 1.48f, 1.48f};
  const float n_ni[AER_NWL] = {8.7e-7f, 3e-8f, 1.2e-8f, 1.9e-4f, 0.12f, 0.16f, 0.14f, 0.04f, 0.02f, 0.014f, 0.017f, 0.03f, 0.012f, 0.005f, 0.006f, 0.003f, 0.002f, 0.0018f, 0.0014f, 0.002f};
  for (int w = 0; w < AER_NWL; w++) {
    // classes: 0 so4, 1 no3, 2 cl, 3 nh4, 4 na, 5 oin, 6 oc, 7 bc, 8 water
    for (int c : {0, 1, 3}) { nr[c][w] = s_nr[w]; ni[c][w] = s_ni[w]; }
    for (int c : {2, 4}) { nr[c][w] = n_nr[w]; ni[c][w] = n_ni[w]; }
    nr[5][w] = d_nr[w]; ni[5][w] = d_ni[w];
    nr[6][w] = 1.45f; ni[6][w] = w < 4 ? 0.001f : 0.02f;
    nr[7][w] = 1.95f; ni[7][w] = 0.79f;
    nr[8][w] = w_nr[w]; ni[8][w] = w_ni[w];
  }
}

int build_aer_tables(const float nr_in[AER_NCLASS][AER_NWL], const float ni_in[AER_NCLASS][AER_NWL], AerTables &T) {
  // wavelengths (cm): 4 chem SW wavelengths (mieaer: 0.30, 0.40, 0.60, 0.999 um) + RRTMG-LW band centres 1e4/nu (um)
  const double sw_um[4] = {0.30, 0.40, 0.60, 0.999};
  const double lw_nu[16] = {180., 425., 565., 665., 760., 900., 1030., 1130., 1285., 1435., 1640., 1940., 2165., 2315., 2490., 2925.};
  for (int w = 0; w < 4; w++) T.wavelength_cm[w] = sw_um[w] * 1.e-4;
  for (int w = 0; w < 16; w++) T.wavelength_cm[4 + w] = 1.0 / lw_nu[w];
  for (int c = 0; c < AER_NCLASS; c++)
    for (int w = 0; w < AER_NWL; w++) { T.nr[c][w] = nr_in[c][w]; T.ni[c][w] = ni_in[c][w]; }
  T.rmin = 0.005e-4; T.rmax = 50.e-4;
  const double xrmin = std::log(T.rmin), xrmax = std::log(T.rmax);
  T.coef.assign((size_t)AER_NWL * AER_NQ * AER_NREFR * AER_NREFI * AER_NCOEF_PAD, 0.f);
  const int nsiz = AER_NSIZ;
  std::vector<double> rs(nsiz), f[AER_NQ];
  for (int q = 0; q < AER_NQ; q++) f[q].resize(nsiz);
  for (int n = 0; n < nsiz; n++) {
    const double xr = std::cos(M_PI * (n + 0.5) / nsiz);
    rs[n] = std::exp(0.5 * (xr * (xrmax - xrmin) + xrmax + xrmin));
  }
  for (int w = 0; w < AER_NWL; w++) {
    // refractive-index grid of this wavelength: spans the species values (mieaer: refrmin..refrmax linear, refimin..refimax geometric)
    double rmin_ = 1e9, rmax_ = -1e9, imin_ = 1e9, imax_ = -1e9;
    for (int c = 0; c < AER_NCLASS; c++) {
      rmin_ = std::min(rmin_, (double)T.nr[c][w]); rmax_ = std::max(rmax_, (double)T.nr[c][w]);
      imin_ = std::min(imin_, (double)T.ni[c][w]); imax_ = std::max(imax_, (double)T.ni[c][w]);
    }
    imin_ = std::max(imin_, 1.e-9);
    imax_ = std::max(imax_, imin_ * 10.0);
    if (rmax_ - rmin_ < 1e-3) rmax_ = rmin_ + 1e-3;
    T.refr_lo[w] = (float)rmin_; T.refr_hi[w] = (float)rmax_; T.refi_lo[w] = (float)imin_; T.refi_hi[w] = (float)imax_;
    for (int ir = 0; ir < AER_NREFR; ir++) {
      const double refr = rmin_ + (rmax_ - rmin_) * ir / (AER_NREFR - 1);
      for (int ii = 0; ii < AER_NREFI; ii++) {
        const double refi = imin_ * std::pow(imax_ / imin_, (double)ii / (AER_NREFI - 1));
        for (int n = 0; n < nsiz; n++) {
          const double x = 2.0 * M_PI * rs[n] / T.wavelength_cm[w];
          double qe, qs, g;
          mie_sphere(x, refr, refi, qe, qs, g);
          f[0][n] = std::log(std::max(qe, 1e-300));
          f[1][n] = std::log(std::max(qs, 1e-300));
          f[2][n] = std::log(std::max(g, 1e-6));
        }
        for (int q = 0; q < AER_NQ; q++) {
          float *c = &T.coef[((((size_t)w * AER_NQ + q) * AER_NREFR + ir) * AER_NREFI + ii) * AER_NCOEF_PAD];
          for (int j = 0; j < AER_NCOEF; j++) {          // chebft
            double s = 0.0;
            for (int k = 0; k < nsiz; k++) s += f[q][k] * std::cos(M_PI * j * (k + 0.5) / nsiz);
            c[j] = (float)(2.0 * s / nsiz);
          }
        }
      }
    }
  }
  return 0;
}

}  // namespace arc
