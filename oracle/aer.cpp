// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.hpp).
// Aerosol optical properties: CPU restatement, in double precision and written independently of the product code, of the
// published algorithm of WRF-Chem v3.9.1 chem/module_optical_averaging.F (optical_prep_sectional / optical_prep_modal /
// mieaer: Fast et al. 2006; Barnard et al. 2010; Ghan & Zaveri 2007).
//
// PARITY UNPINNED: that Fortran module is not part of the reference repository and cannot be
// consulted here; this oracle pins the CUDA stage against an independent implementation of the same description and the
// tests additionally compare the Chebyshev tables with direct Mie theory.  Species refractive indices are supplied by the
// caller (the product's defaults are passed in by the tests) so both sides use identical physical data.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <string>
#include <vector>

#include "../include/arc_rad.h"
#include "adapter_common.hpp"

namespace {

const int NWL = 20, NCLS = 9, NRE = 7, NIM = 7, NCO = 50, NSZ = 200;
const double DENS[NCLS] = {1.8, 1.8, 2.2, 1.8, 2.2, 2.6, 1.0, 1.7, 1.0};

struct OrcAer {
  bool ready = false;
  double lam[NWL], nr[NCLS][NWL], ni[NCLS][NWL], rlo[NWL], rhi[NWL], ilo[NWL], ihi[NWL];
  std::vector<double> coef;    // [wl][q][ir][ii][NCO]
  double lrmin, lrmax;
};
OrcAer A;
std::string g_aerr;

// Lorenz-Mie efficiencies (Bohren & Huffman appendix A formulation, m = n + ik)
void mie(double x, std::complex<double> m, double &qe, double &qs, double &g) {
  using C = std::complex<double>;
  const C mx = m * x;
  const int nstop = int(x + 4.0 * std::pow(x, 1.0 / 3.0) + 2.0);
  const int nmax = int(std::max(double(nstop), std::abs(mx))) + 15;
  std::vector<C> d(nmax + 1, C(0, 0));
  for (int n = nmax; n > 0; --n) d[n - 1] = double(n) / mx - 1.0 / (d[n] + double(n) / mx);
  double pm1 = std::cos(x), p0 = std::sin(x), cm1 = -std::sin(x), c0 = std::cos(x);
  C a_prev, b_prev;
  double se = 0, ss = 0, sg = 0;
  for (int n = 1; n <= nstop; ++n) {
    const double p1 = (2 * n - 1) / x * p0 - pm1, c1 = (2 * n - 1) / x * c0 - cm1;
    const C xi1(p1, -c1), xi0(p0, -c0);
    const C ta = d[n] / m + double(n) / x, tb = d[n] * m + double(n) / x;
    const C a = (ta * p1 - p0) / (ta * xi1 - xi0), b = (tb * p1 - p0) / (tb * xi1 - xi0);
    se += (2 * n + 1) * (a.real() + b.real());
    ss += (2 * n + 1) * (std::norm(a) + std::norm(b));
    sg += (2.0 * n + 1.0) / (n * (n + 1.0)) * (a * std::conj(b)).real();
    if (n > 1) sg += (n - 1.0) * (n + 1.0) / n * ((a_prev * std::conj(a)).real() + (b_prev * std::conj(b)).real());
    a_prev = a; b_prev = b; pm1 = p0; p0 = p1; cm1 = c0; c0 = c1;
  }
  qe = 2 * se / (x * x); qs = 2 * ss / (x * x); g = ss > 0 ? 2 * sg / ss : 0;
}

void build(const float *nr, const float *ni) {
  const double sw[4] = {0.30e-4, 0.40e-4, 0.60e-4, 0.999e-4};
  const double nu[16] = {180, 425, 565, 665, 760, 900, 1030, 1130, 1285, 1435, 1640, 1940, 2165, 2315, 2490, 2925};
  for (int w = 0; w < NWL; w++) A.lam[w] = w < 4 ? sw[w] : 1.0 / nu[w - 4];
  for (int c = 0; c < NCLS; c++) for (int w = 0; w < NWL; w++) { A.nr[c][w] = nr[c * NWL + w]; A.ni[c][w] = ni[c * NWL + w]; }
  A.lrmin = std::log(0.005e-4); A.lrmax = std::log(50e-4);
  A.coef.assign((size_t)NWL * 3 * NRE * NIM * NCO, 0.0);
  std::vector<double> r(NSZ), y[3];
  for (auto &v : y) v.resize(NSZ);
  for (int k = 0; k < NSZ; k++) r[k] = std::exp(0.5 * (std::cos(M_PI * (k + 0.5) / NSZ) * (A.lrmax - A.lrmin) + A.lrmax + A.lrmin));
  for (int w = 0; w < NWL; w++) {
    double a = 1e9, b = -1e9, c = 1e9, d = -1e9;
    for (int s = 0; s < NCLS; s++) { a = std::min(a, A.nr[s][w]); b = std::max(b, A.nr[s][w]); c = std::min(c, A.ni[s][w]); d = std::max(d, A.ni[s][w]); }
    c = std::max(c, 1e-9); d = std::max(d, 10 * c); if (b - a < 1e-3) b = a + 1e-3;
    // the product stores its grid bounds in single precision: use the same rounded bounds so both sides interpolate alike
    A.rlo[w] = (float)a; A.rhi[w] = (float)b; A.ilo[w] = (float)c; A.ihi[w] = (float)d;
    for (int ir = 0; ir < NRE; ir++) for (int ii = 0; ii < NIM; ii++) {
      const double re = a + (b - a) * ir / (NRE - 1), im = c * std::pow(d / c, double(ii) / (NIM - 1));
      for (int k = 0; k < NSZ; k++) {
        double qe, qs, g; mie(2 * M_PI * r[k] / A.lam[w], {re, im}, qe, qs, g);
        y[0][k] = std::log(std::max(qe, 1e-300)); y[1][k] = std::log(std::max(qs, 1e-300)); y[2][k] = std::log(std::max(g, 1e-6));
      }
      for (int q = 0; q < 3; q++) for (int j = 0; j < NCO; j++) {
        double s = 0; for (int k = 0; k < NSZ; k++) s += y[q][k] * std::cos(M_PI * j * (k + 0.5) / NSZ);
        A.coef[((((size_t)w * 3 + q) * NRE + ir) * NIM + ii) * NCO + j] = 2 * s / NSZ;
      }
    }
  }
  A.ready = true;
}

void sphere(int w, double r, double re, double im, double &pe, double &ps, double &pg) {
  r = std::min(std::max(r, std::exp(A.lrmin)), std::exp(A.lrmax));
  double tr = std::min(std::max((re - A.rlo[w]) / (A.rhi[w] - A.rlo[w]) * (NRE - 1), 0.0), double(NRE - 1));
  int ir = std::min(int(tr), NRE - 2); const double t = tr - ir;
  double ti = std::min(std::max((std::log(std::max(im, 1e-30)) - std::log(A.ilo[w])) / (std::log(A.ihi[w]) - std::log(A.ilo[w])) * (NIM - 1), 0.0), double(NIM - 1));
  int ii = std::min(int(ti), NIM - 2); const double u = ti - ii;
  const double x = (2 * std::log(r) - A.lrmax - A.lrmin) / (A.lrmax - A.lrmin);
  double out[3];
  for (int q = 0; q < 3; q++) {
    // Clenshaw recurrence (a different evaluation order than the product's forward recurrence)
    auto cj = [&](int j) {
      auto C = [&](int a, int b) { return A.coef[((((size_t)w * 3 + q) * NRE + a) * NIM + b) * NCO + j]; };
      return (1 - t) * (1 - u) * C(ir, ii) + t * (1 - u) * C(ir + 1, ii) + (1 - t) * u * C(ir, ii + 1) + t * u * C(ir + 1, ii + 1);
    };
    double b1 = 0, b2 = 0;
    for (int j = NCO - 1; j >= 1; --j) { const double b0 = 2 * x * b1 - b2 + cj(j); b2 = b1; b1 = b0; }
    out[q] = std::exp(x * b1 - b2 + 0.5 * cj(0));
  }
  pe = out[0]; ps = std::min(out[1], out[0]); pg = out[2];
}

}  // namespace

extern "C" {

const char *arc_oracle_aer_last_error(void) { return g_aerr.c_str(); }

int arc_oracle_aer_init(const float *refr, const float *refi) {
  if (!refr || !refi) { g_aerr = "oracle aer: refractive indices required"; return ARC_ERR_BAD_ARG; }
  build(refr, refi);
  return 0;
}

int arc_oracle_aer_optics(const ArcDims *d, const ArcAerIn *in, ArcAerOut *out) {
  if (!A.ready) { g_aerr = "oracle aer: not initialised"; return ARC_ERR_NOT_INIT; }
  const orc::Idx ix(*d);
  const int nsec = in->mode == ARC_AER_SECTIONAL ? in->nbin : 8;
  const double lo = 3.90625e-6, hi = 1.0e-3;
  for (int j = d->jts; j <= d->jte; j++) for (int k = d->kts; k <= d->kte; k++) for (int i = d->its; i <= d->ite; i++) {
    const size_t q = ix.at3(i, k, j);
    const double rho = 1.0 / in->alt[q];
    std::vector<std::vector<double>> vol(nsec, std::vector<double>(NCLS, 0.0));
    std::vector<double> num(nsec, 0.0);
    if (in->mode == ARC_AER_SECTIONAL) {
      for (int s = 0; s < nsec; s++) {
        for (int m = 0; m < in->nspec[s]; m++) vol[s][in->cls[s][m]] += std::max((double)in->mass[s][m][q], 0.0) * rho * 1e-12 / DENS[in->cls[s][m]];
        num[s] = std::max((double)in->num[s][q], 0.0) * rho * 1e-6;
      }
    } else {
      for (int md = 0; md < in->nbin; md++) {
        std::vector<double> vm(NCLS, 0.0); double vt = 0;
        for (int m = 0; m < in->nspec[md]; m++) { const double v = std::max((double)in->mass[md][m][q], 0.0) * rho * 1e-12 / DENS[in->cls[md][m]]; vm[in->cls[md][m]] += v; vt += v; }
        const double nm = std::max((double)in->num[md][q], 0.0) * rho * 1e-6;
        if (!(vt > 1e-30) || !(nm > 1e-20)) continue;
        const double ls = std::log((double)in->sigmag[md]);
        const double dgn = std::cbrt(vt / (M_PI / 6 * nm)) * std::exp(-1.5 * ls * ls), dgv = dgn * std::exp(3 * ls * ls);
        double fnp = 0, fvp = 0;
        for (int s = 0; s < nsec; s++) {
          const double dhi = lo * std::exp(std::log(hi / lo) * (s + 1) / nsec);
          const double fn = s == nsec - 1 ? 1.0 : 0.5 * (1 + std::erf(std::log(dhi / dgn) / (std::sqrt(2.0) * ls)));
          const double fv = s == nsec - 1 ? 1.0 : 0.5 * (1 + std::erf(std::log(dhi / dgv) / (std::sqrt(2.0) * ls)));
          num[s] += nm * (fn - fnp);
          for (int c = 0; c < NCLS; c++) vol[s][c] += vm[c] * (fv - fvp);
          fnp = fn; fvp = fv;
        }
      }
    }
    double ext[NWL] = {0}, sca[NWL] = {0}, gsc[NWL] = {0};
    for (int s = 0; s < nsec; s++) {
      double vdry = 0; for (int c = 0; c < NCLS - 1; c++) vdry += vol[s][c];
      if (!(vdry > 1e-30)) continue;
      const double vwet = vdry + vol[s][NCLS - 1];
      const double dlo = lo * std::exp(std::log(hi / lo) * s / nsec), dhi = lo * std::exp(std::log(hi / lo) * (s + 1) / nsec);
      double n = num[s];
      double dp = n > 1e-20 ? std::cbrt(6 / M_PI * vdry / n) : 0.0;
      if (!(dp >= dlo && dp <= dhi)) { dp = std::sqrt(dlo * dhi); n = 6 / M_PI * vdry / (dp * dp * dp); }
      const double r = 0.5 * dp * std::cbrt(vwet / vdry);
      const double wgt = n * M_PI * r * r;
      for (int w = 0; w < NWL; w++) {
        double re = 0, im = 0;
        for (int c = 0; c < NCLS; c++) { re += vol[s][c] / vwet * A.nr[c][w]; im += vol[s][c] / vwet * A.ni[c][w]; }
        double pe, ps, pg; sphere(w, r, re, im, pe, ps, pg);
        ext[w] += wgt * pe; sca[w] += wgt * ps; gsc[w] += wgt * ps * pg;
      }
    }
    const double dzcm = in->dz8w[q] * 100.0;
    for (int w = 0; w < 4; w++) {
      out->tauaer[w][q] = (float)(ext[w] * dzcm);
      out->waer[w][q] = (float)(ext[w] > 0 ? sca[w] / ext[w] : 1.0);
      out->gaer[w][q] = (float)(sca[w] > 0 ? gsc[w] / sca[w] : 0.0);
    }
    for (int w = 0; w < 16; w++) {
      const double ab = std::max(ext[4 + w] - sca[4 + w], 0.0);
      out->tauaerlw[w][q] = (float)(ab * dzcm);
      if (out->extaerlw[w]) out->extaerlw[w][q] = (float)(ab * 1e5);
    }
  }
  return 0;
}

}  // extern "C"
