// Aerosol optical-property stage: argument blocks shared by aer_optics.cu and api.cu.
#pragma once
#include <string>

#include "aer_tables.h"
#include "args.h"

namespace arc {

struct AerDev {
  const float *coef;                       // [wl][q][refr][refi][52]
  float nr[AER_NCLASS][AER_NWL], ni[AER_NCLASS][AER_NWL];
  float refr_lo[AER_NWL], refr_hi[AER_NWL], lnrefi_lo[AER_NWL], lnrefi_hi[AER_NWL];
  float xrmin, xrmax, rmin, rmax;
};

struct AerSpecList {                      // device copy of the species lists
  int mode, nbin;
  int nspec[AER_MAXBIN];
  int cls[AER_MAXBIN][AER_MAXSPEC];
  const float *mass[AER_MAXBIN][AER_MAXSPEC];
  const float *num[AER_MAXBIN];
  float sigmag[AER_MAXBIN];
};

struct AerArgs {
  Geo geo;
  int npts;                                // tile columns x levels
  int nsec;                                // output sections
  const AerSpecList *sl;                   // device
  const float *alt, *dz8w;
  float *ws;                               // [nsec][AER_WS_N][npts]
  float *tauaer[4], *gaer[4], *waer[4], *tauaerlw[16], *extaerlw[16];
};

int aer_init(const float *nr, const float *ni, std::string &err);
bool aer_ready();
void aer_finalize();
const AerTables &aer_tables();
int aer_run(AerArgs &a, const AerSpecList &sl, cudaStream_t s, std::string &err);

}  // namespace arc
