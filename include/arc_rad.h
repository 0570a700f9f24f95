/*
 * arc_rad.h -- C ABI of the B200-native RRTMG SW/LW radiation hot path
 * (drop-in for the WRF-Chem ARC "clean atmosphere" patched RRTMG of
 *  douglowe/WRFChem-ARC-Interactions, v3.9.1).
 *
 * Every entry point replaces one Fortran entry point of the reference; the
 * struct members carry exactly the reference's dummy arguments (same names,
 * same units, same WRF memory order).  Citations are into
 * WRF-Chem_code/v3.9.1/phys/ of the reference:
 *
 *   arc_rad_init      <- rrtmg_swinit  module_ra_rrtmg_sw.F:11211-11236
 *                        rrtmg_lwinit  module_ra_rrtmg_lw.F:12845-12876
 *   arc_rad_sw        <- RRTMG_SWRAD   module_ra_rrtmg_sw.F:9901-11207
 *   arc_rad_lw        <- RRTMG_LWRAD   module_ra_rrtmg_lw.F:11451-12700
 *   arc_rad_finalize  <- (nothing; the reference never frees its module tables)
 *   arc_rad_last_error<- wrf_error_fatal / stop messages (SW: 40 sites, LW: 30)
 *
 * Array conventions (module_radiation_driver.F:1528-1577, 1944-2001):
 *   3-D  REAL(4) (ims:ime, kms:kme, jms:jme)   i contiguous, then k, then j
 *   2-D  REAL(4) (ims:ime, jms:jme)
 *   flux profiles (ims:ime, kms:kme+2, jms:jme)
 *   4-D  (ims:ime, kms:kme, jms:jme, n)
 * Only the tile its:ite, kts:kte, jts:jte is read / written; halo cells are
 * never touched.  Each OPTIONAL Fortran dummy is a nullable pointer; each F_Qx
 * LOGICAL is an int flag.  The caller owns every array; nothing is retained
 * after return.
 *
 * New relative to the reference: `memspace` (pointers are host or device
 * memory) and `variant_mask`, which hands the whole tile over once with a
 * call-variant axis instead of the reference's two internal solver calls.
 *
 * Streams (memspace = ARC_MEM_DEVICE).  The library enqueues its kernels on
 * private non-blocking CUDA streams (the main one is arc_rad_stream()) and does
 * NOT order them against any stream of the caller.  Contract: every input array,
 * and every INOUT output array the caller pre-filled, must be complete before
 * the call - cudaStreamSynchronize / cudaDeviceSynchronize on the producing
 * streams, or an event the producer recorded that the caller has waited on with
 * cudaStreamWaitEvent on arc_rad_stream() BEFORE calling.  On return every output
 * is complete: each entry point synchronises its own streams before it returns.
 * With ARC_MEM_HOST the call is synchronous in the ordinary host sense.
 */
#ifndef ARC_RAD_H
#define ARC_RAD_H

#ifdef __cplusplus
extern "C" {
#endif

#define ARC_OK                 0
#define ARC_ERR_NOT_INIT       1
#define ARC_ERR_BAD_ARG        2
#define ARC_ERR_IO             3   /* table file missing / malformed record        */
#define ARC_ERR_MISSING_FIELD  4   /* 'missing fields required for aerosol radiation' SW:10288-10305 */
#define ARC_ERR_NEG_AOD        5   /* 'Negative total optical depth'  SW:11031, LW:12613 */
#define ARC_ERR_RADIUS         6   /* effective radius / dge out of table bounds SW:2164,2281 LW:2844,2893 */
#define ARC_ERR_CUDA           7
#define ARC_ERR_CONFIG         8   /* clean_atm_diag>0 needs aer_ra_feedback>0, chemics_init.F:406-408 */
#define ARC_ERR_UNSUPPORTED    9

#define ARC_MEM_HOST   0
#define ARC_MEM_DEVICE 1

/* call variants (Ghan 2012 decomposition streams) */
#define ARC_VAR_FULL       1   /* clouds + aerosol          -> SWUPT ...        */
#define ARC_VAR_CLEAR      2   /* no clouds, with aerosol   -> SWUPTC ...       */
#define ARC_VAR_CLEAN      4   /* clouds, no aerosol        -> SWUPTCLN ...     */
#define ARC_VAR_CLEANCLEAR 8   /* no clouds, no aerosol (computed and discarded by the reference,
                                  SW:9466, LW:11026; exposed here as an option) */

/* the 18 WRF index integers, in the reference's order */
typedef struct ArcDims {
  int ids, ide, jds, jde, kds, kde;
  int ims, ime, jms, jme, kms, kme;
  int its, ite, jts, jte, kts, kte;
} ArcDims;

/* init-time configuration (rrtmg_swinit/rrtmg_lwinit arguments + module_model_constants) */
typedef struct ArcConfig {
  float cp;        /* specific heat of dry air, module_model_constants cp = 7*287/2 = 1004.5 */
  float p_top;     /* model top pressure (Pa); LW nlayers = kme + nint(p_top*0.01/4) - 1  LW:12861 */
  int   kme;       /* memory upper k bound used for nlayers                                 */
  int   device;    /* CUDA device ordinal (-1: current)                                     */
  const char *inline_tables; /* path of rrtmg_inline_tables.bin (NULL: $ARC_RAD_TABLES or package default) */
} ArcConfig;

/* --------------------------------------------------------------------------------------------
 * RRTMG_SWRAD  (SW:9901-9945 argument list; declarations SW:9949-10102)                        */
typedef struct ArcSwIn {
  int memspace;            /* ARC_MEM_HOST / ARC_MEM_DEVICE */
  int variant_mask;        /* 0: FULL|CLEAR (+CLEAN when clean_atm_diag>0)                     */
  /* scalars */
  float radt, degrad, declin, solcon, xtime, gmt, r, g, julian;
  int julday, icloud, warm_rain, is_cammgmp_used;
  int has_reqc, has_reqi, has_reqs;
  int o3input, aer_opt, no_src, sf_surface_physics, mp_physics;
  int aer_ra_feedback, progn, clean_atm_diag;
  int f_qv, f_qc, f_qr, f_qi, f_qs, f_qg, f_qndrop;   /* -1 = argument not present */
  /* 3-D (i,k,j) */
  const float *t3d, *t8w, *p3d, *p8w, *pi3d, *rho3d, *dz8w;
  const float *cldfra3d, *lradius, *iradius;
  const float *qv3d, *qc3d, *qr3d, *qi3d, *qs3d, *qg3d, *qndrop3d;
  const float *o33d;
  const float *re_cloud, *re_ice, *re_snow;
  const float *f_ice_phy, *f_rain_phy;
  const float *tauaer300, *tauaer400, *tauaer600, *tauaer999;
  const float *gaer300, *gaer400, *gaer600, *gaer999;
  const float *waer300, *waer400, *waer600, *waer999;
  const float *aerod;                                   /* (i,k,j,no_src >= 6), aer_opt=1: the six ECMWF aerosol types (iaer = 6); not with clean_atm_diag */
  const float *tauaer3d_sw, *ssaaer3d_sw, *asyaer3d_sw; /* (i,k,j,14) pointers, aer_opt=2,3  */
  /* 2-D (i,j) */
  const float *xcoszen, *albedo, *tsk, *xland, *xice, *snow;
  const float *alswvisdir, *alswvisdif, *alswnirdir, *alswnirdif;
  const float *xlat, *xlong;                            /* unused by the reference body */
} ArcSwIn;

typedef struct ArcSwOut {
  float *rthratensw;                                    /* (i,k,j) written only where coszen>0 */
  float *gsw, *swcf, *coszr;                            /* (i,j) */
  float *swupt, *swuptc, *swuptcln, *swdnt, *swdntc, *swdntcln;
  float *swupb, *swupbc, *swupbcln, *swdnb, *swdnbc, *swdnbcln;
  float *swvisdir, *swvisdif, *swnirdir, *swnirdif;
  float *swddir, *swddni, *swddif;
  float *swupflx, *swupflxc, *swupflxcln;               /* (i, kms:kme+2, j), optional */
  float *swdnflx, *swdnflxc, *swdnflxcln;
  /* extension: the 4th (clean + clear) stream, TOA/surface, optional */
  float *swuptclnc, *swdntclnc, *swupbclnc, *swdnbclnc;
} ArcSwOut;

/* --------------------------------------------------------------------------------------------
 * RRTMG_LWRAD  (LW:11451-11486 argument list; declarations LW:11493-11604)                     */
typedef struct ArcLwIn {
  int memspace;
  int variant_mask;
  float r, g, julian;
  int yr;
  int icloud, warm_rain, is_cammgmp_used;
  int has_reqc, has_reqi, has_reqs;
  int o3input, mp_physics;
  int aer_ra_feedback, progn, clean_atm_diag;
  int f_qv, f_qc, f_qr, f_qi, f_qs, f_qg, f_qndrop;
  const float *p8w, *p3d, *pi3d, *dz8w, *t3d, *t8w, *rho3d;
  const float *cldfra3d, *lradius, *iradius;
  const float *qv3d, *qc3d, *qr3d, *qi3d, *qs3d, *qg3d, *qndrop3d;
  const float *o33d;
  const float *re_cloud, *re_ice, *re_snow;
  const float *f_ice_phy, *f_rain_phy;
  const float *tauaerlw[16];                            /* tauaerlw1 .. tauaerlw16 */
  const float *emiss, *tsk, *xland, *xice, *snow;       /* (i,j) */
} ArcLwIn;

typedef struct ArcLwOut {
  float *rthratenlw;                                    /* (i,k,j) */
  float *glw, *olr, *lwcf;                              /* (i,j) */
  float *lwupt, *lwuptc, *lwuptcln, *lwdnt, *lwdntc, *lwdntcln;
  float *lwupb, *lwupbc, *lwupbcln, *lwdnb, *lwdnbc, *lwdnbcln;
  float *lwupflx, *lwupflxc, *lwupflxcln;               /* (i, kms:kme+2, j), optional */
  float *lwdnflx, *lwdnflxc, *lwdnflxcln;
  float *lwuptclnc, *lwdntclnc, *lwupbclnc, *lwdnbclnc; /* extension, optional */
} ArcLwOut;

/* --------------------------------------------------------------------------------------------
 * Optional intermediate taps used by the parity tests (all pointers nullable, always HOST
 * memory).  Column index c = (j-jts)*(ite-its+1) + (i-its).  Layer index 0 = surface.
 * nlay = kte-kts+2 (SW) or nlayers (LW).                                                      */
typedef struct ArcDebug {
  int   *laytrop;      /* [ncol]                                  */
  int   *jp, *jt, *jt1, *indfor, *indself;   /* [ncol][nlay]       */
  int   *indminor;     /* LW only [ncol][nlay]                    */
  float *fac00, *fac01, *fac10, *fac11;      /* [ncol][nlay]       */
  unsigned char *cldmask;  /* McICA mask [ncol][nlay][ngpt] 0/1     */
  float *taug;         /* [ncol][nlay][ngpt]                      */
  float *taur;         /* SW [ncol][nlay][ngpt]; LW: fracs         */
  float *sfluxzen;     /* SW [ncol][ngpt]                         */
  float *taucmc;       /* [ncol][nlay][ngpt] cloud optical depth (SW delta-scaled) */
  float *hr;           /* heating rate K/day [ncol][nlay], full    */
  float *sw_cond;      /* SW, filled by the oracle only [ncol]: min over (g, layer, stream) of |1 - (k*mu0)^2| in
                          reftra_sw.  At k*mu0 = 1 the reference's two-stream solution is a table-quantised 0/0
                          (SW:2629-2660): its own output there is rounding noise, so parity tests exclude columns
                          whose conditioning is below a stated threshold and report how many. */
} ArcDebug;

int  arc_rad_init(const ArcConfig *cfg, const char *sw_data_path, const char *lw_data_path);
int  arc_rad_sw(const ArcDims *d, const ArcSwIn *in, ArcSwOut *out);
int  arc_rad_lw(const ArcDims *d, const ArcLwIn *in, ArcLwOut *out);
/* One radiation step: RRTMG_LWRAD then RRTMG_SWRAD on the same tile (the order of radiation_driver, DRV:1526-2009).  With
 * host arrays both adapters run inside one j-slab pipeline (shared inputs uploaded once); results equal the two calls. */
int  arc_rad_lwsw(const ArcDims *d, const ArcLwIn *lwin, ArcLwOut *lwout, const ArcSwIn *swin, ArcSwOut *swout);
/* same as above with intermediate taps (tests only) */
int  arc_rad_sw_debug(const ArcDims *d, const ArcSwIn *in, ArcSwOut *out, ArcDebug *dbg);
int  arc_rad_lw_debug(const ArcDims *d, const ArcLwIn *in, ArcLwOut *out, ArcDebug *dbg);
void arc_rad_finalize(void);
const char *arc_rad_last_error(void);
int  arc_rad_lw_nlayers(void);                 /* LW:12861 value fixed at init */

/* bookkeeping done by radiation_driver around the two calls
 * (module_radiation_driver.F:1692-1702, 2180-2194, 2308-2377):
 *   rthratenlw = rthraten(after LW); rthraten += sw; rthratensw = rthraten - rthratenlw;
 *   swdown = gsw/(1-albedo); and time accumulation ac* += flux*dt.                          */
int  arc_rad_driver_post(const ArcDims *d, int memspace,
                         const float *rthratenlw, const float *rthratensw, float *rthraten,
                         const float *gsw, const float *albedo, float *swdown);

/* Solar geometry of radiation_driver (module_radiation_driver.F:2595-2666, called at DRV:954-973):
 *   radconst:    declination and solar constant 1370*eccfac(julian)        (host scalar arithmetic, FP32 like the reference)
 *   calc_coszen: cos(zenith) and hour angle with the equation of time, at the time the caller passes (xtime + radt/2)   */
void arc_rad_radconst(float xtime, float julian, float degrad, float dpd, float *declin, float *solcon);
int  arc_rad_calc_coszen(const ArcDims *d, int memspace, float julian, float xtime, float gmt, float declin, float degrad,
                         const float *xlon, const float *xlat, float *coszen, float *hrang);
/* time accumulation AC{SW,LW}{UP,DN}{T,B}[C] += flux*DTaccum for `nfields` 2-D field pairs (DRV:2308-2377; the reference
 * does not accumulate the *CLN fields) */
int  arc_rad_accumulate(const ArcDims *d, int memspace, float dtaccum, int nfields, const float *const *flux, float *const *acc);

/* Domain statistics of `nfields` 2-D (i,j) fields over the tile: out[f][5] = {sum, sum of squares, count, min, max}
 * in double precision; `fields` is a host array of pointers in `memspace`, `out` lives in `memspace` too.
 * Replaces calc_standard_stats' mean/SD/SE inputs (analysis_scripts/NCL_extraction_package/misc_stats_library.ncl:396-461);
 * partial results of several GPUs combine with one sum- and one max-all-reduce. */
int  arc_rad_domain_stats(const ArcDims *d, int memspace, int nfields, const float *const *fields, double *out);
/* Moran's I of `nfields` 2-D (i,j) fields over the tile, out[f] in single precision, exactly as calc_morans_i_2D evaluates it
 * for calc_standard_stats (neighbour + manhattan options: ordered pairs of edge-sharing cells, N-1 variance;
 * misc_stats_library.ncl:196-371).  corrected SE = SE * I (misc_stats_library.ncl:449). */
int  arc_rad_morans_i(const ArcDims *d, int memspace, int nfields, const float *const *fields, float *out);
/* Counts of the reference's aerosol warnings in the most recent call (or LW + SW pair): (column, band) pairs whose chem-aerosol
 * column optical depth exceeded 6 in the shortwave and was rescaled to 6 (SW:11034-11069), or exceeded 5 in the longwave (warning
 * only, LW:12616-12627).  Either pointer may be NULL. */
void arc_rad_warning_counts(int *sw_aod_capped, int *lw_aod_large);
/* cal_cldfra1 (module_radiation_driver.F:2886-3122; radiation_driver calls it for icloud = 1, DRV:1104-1118): cloud fraction
 * CLDFRA(i,k,j) from QV, QC, QI, QS, T, p, tile levels kts..kte.  f_q*: 1 = .TRUE., 0 = .FALSE., < 0 = argument not PRESENT;
 * f_ice_phy and cldfra1_flag (INTEGER(i,k,j): 1 no condensate, 2 saturated, 3 partial) may be NULL.  Bit-exact with the
 * reference's arithmetic (unfused, glibc EXP and **). */
int  arc_rad_cal_cldfra1(const ArcDims *d, int memspace, const float *qv, const float *qc, const float *qi, const float *qs, int f_qv,
                         int f_qc, int f_qi, int f_qs, const float *t_phy, const float *p_phy, const float *f_ice_phy, int mp_physics,
                         float *cldfra, int *cldfra1_flag);
/* cal_cldfra2 (module_radiation_driver.F:2801-2874; icloud = 2, DRV:1205): CLDFRA = 1 where QC + QI > 1e-6 (QC alone when
 * F_QI is false; 0 everywhere when F_QC is false), tile levels kts..kte. */
int  arc_rad_cal_cldfra2(const ArcDims *d, int memspace, const float *qc, const float *qi, int f_qc, int f_qi, float *cldfra);
/* cal_cldfra3 (module_radiation_driver.F:3140-3274 with find_cloudLayers / adjust_cloud* :3281-3599; icloud = 3, DRV:1228):
 * CLDFRA(i,k,j) from relative humidity against a grid-size (gridkm) and land / ocean (XLAND) dependent threshold, fractional
 * clouds removed above the diagnosed tropopause and in the well-mixed layer, then per cloud layer a made-up condensate
 * profile ADDED to qc / qi (INOUT, as in the reference; the caller saves and restores qs, DRV:1217-1225).  rslf / rsif come from
 * module_mp_thompson, which the reference repository does not contain: the published Flatau polynomials are used (unpinned). */
int  arc_rad_cal_cldfra3(const ArcDims *d, int memspace, float *cldfra, const float *qv, float *qc, float *qi, const float *qs,
                         const float *p, const float *t, const float *rho, const float *xland, float gridkm);
/* ozn_time_int (module_radiation_driver.F:3993-4098; o3input = 2, DRV:1250): ozmixm(ims:ime, levsiz, jms:jme, num_months) ->
 * ozmixt(ims:ime, levsiz, jms:jme), linear in time between the mid-month days that bracket JULIAN + 1 (December-January wraps). */
int  arc_rad_ozn_time_int(const ArcDims *d, int memspace, int julday, float julian, int levsiz, int num_months, const float *ozmixm,
                          float *ozmixt);
/* ozn_p_int (module_radiation_driver.F:4100-4234; DRV:1256): ozmixt on the data levels pin(levsiz) (HOST array in either
 * memspace; Pa, top down, strictly increasing, levsiz <= 128) -> o3vmr(i,k,j) at the model pressures p(i,k,j): linear in
 * pressure inside the data range, scaled by p / pin(1) above it, held below it.  kts must be 1.  Bit-exact (unfused, IEEE /). */
int  arc_rad_ozn_p_int(const ArcDims *d, int memspace, const float *p, const float *pin, int levsiz, const float *ozmixt, float *o3vmr);
/* aer_time_int / aer_p_int (module_radiation_driver.F:4236-4343, 4345-4506; aer_opt = 1): the monthly Tegen aerosol climatology
 * aerodm(ims:ime, levsiz, jms:jme, num_months, no_src) interpolated in time (as ozn_time_int, per aerosol type) and then to the model
 * pressures: AEROD(i,k,j,1:no_src) = interpolated value x (pf(k) - pf(k+1)), TOTAOD(i,j) = their sum over types and levels.  `pin`
 * is a HOST array (hPa, top down, strictly increasing); p and pf (= p8w) in Pa; kts must be 1.  Bit-exact. */
int  arc_rad_aer_time_int(const ArcDims *d, int memspace, int julday, float julian, int levsiz, int num_months, int no_src,
                          const float *aerodm, float *aerodt);
int  arc_rad_aer_p_int(const ArcDims *d, int memspace, const float *p, const float *pin, int levsiz, const float *aerodt, float *aerod,
                       int no_src, const float *pf, float *totaod);
/* Order statistics of `nfields` 2-D (i,j) fields over the tile: out[f * nperc + q] = sorted(field f)[round(0.01 * perc[q] * (N - 1))],
 * the element calc_boxplot_stats picks (misc_stats_library.ncl:145-189); perc = {50, 25, 75, 5, 95} gives calc_standard_stats'
 * median, lower / upper quartile, 5th / 95th percentile (ncl:439-445).  Exact (radix selection, no interpolation).  The 5-cell
 * domain trim of calculate_domain_stats (data_extraction_library.ncl:318-322) is a matter of the tile bounds in `d`.
 * With several GPUs the field must be gathered first (an order statistic does not combine from partial results). */
int  arc_rad_percentiles(const ArcDims *d, int memspace, int nfields, const float *const *fields, int nperc, const float *perc, float *out);
/* Host-only probe (no GPU needed): parse and g-point-reduce the table files as arc_rad_init does; returns the element count
 * of the reduced table `name` ("sw16.absa", "lw3.ka_mn2o", "lw_nlayers" ...) and copies it to buf when cap suffices;
 * name == NULL only validates the files.  Negative return = -ARC_ERR_*.  (sw_kgbNN / cmbgbNN, SW:5022-6065, 11315-12384) */
int  arc_rad_host_table(const char *inline_tables, const char *sw_data_path, const char *lw_data_path, float cp, float p_top,
                        int kme, const char *name, float *buf, int cap);
/* Self-test (GPU): setcoef's jp | jt << 8 | jt1 << 12 for n host (p hPa, T K) pairs (SW:2854-2887), same device code as the prep kernels */
int  arc_rad_selftest_pt(const float *p, const float *t, int n, int *packed);
/* Self-test of the glibc-compatible logf / expf / powf the device path uses for LOG, EXP and ** of the reference
 * (which = 0 logf(x), 1 expf(x), 2 powf(x, y)): out[i] from the host instantiation (on_device = 0; needs no GPU and no init)
 * or the device instantiation (on_device = 1); the caller compares with the C library (SW:2854, 11004-11008; LW:3650). */
int  arc_rad_selftest_libm(int which, const float *x, const float *y, int n, float *out, int on_device);
/* Self-test (GPU): mismatches of the kernels' branch-free division against IEEE division over n random operand pairs */
int  arc_rad_selftest_div(int n, unsigned seed);
/* Self-test (GPU): mismatches of the kernels' reciprocal against the IEEE reciprocal for EVERY float with bit pattern in
 * [lo_bits, hi_bits], both signs (normal range: 0x0D800000 = 2^-100 .. 0x71800000 = 2^100) */
long long arc_rad_selftest_rcp(unsigned lo_bits, unsigned hi_bits);
/* FP32 FMA throughput of the device in TFLOP/s (microbenchmark; roofline denominator of the solver kernels) */
float arc_rad_measure_fp32_tflops(void);

/* --------------------------------------------------------------------------------------------
 * Aerosol optical properties: optical_averaging -> optical_prep_sectional / optical_prep_modal -> mieaer of WRF-Chem
 * v3.9.1 chem/module_optical_averaging.F.  That module is NOT part of the reference repository (only its outputs are
 * consumed: module_radiation_driver.F:113-124, Registry/registry.chem:1332-1390), so this stage restates the published
 * algorithm and is self-consistent only.  Output arrays are exactly the inputs of arc_rad_sw /
 * arc_rad_lw: tauaer300..999, gaer*, waer*, tauaerlw1..16 (+ extaerlw1..16 in 1/km).                               */
#define ARC_AER_MAXBIN   8
#define ARC_AER_MAXSPEC 24
#define ARC_AER_SECTIONAL 1   /* MOSAIC 4 or 8 size sections (aer_op_opt volume_approx)              */
#define ARC_AER_MODAL     2   /* MADE/SORGAM: 3 log-normal modes mapped onto 8 sections              */
/* species classes (density, refractive index): */
enum { ARC_CLS_SO4 = 0, ARC_CLS_NO3, ARC_CLS_CL, ARC_CLS_NH4, ARC_CLS_NA, ARC_CLS_OIN, ARC_CLS_OC, ARC_CLS_BC, ARC_CLS_WATER, ARC_CLS_N };

typedef struct ArcAerIn {
  int memspace;
  int mode;                                  /* ARC_AER_SECTIONAL / ARC_AER_MODAL */
  int nbin;                                  /* sections (4 or 8) or modes (3) */
  int nspec[ARC_AER_MAXBIN];                 /* species per section / mode */
  int cls[ARC_AER_MAXBIN][ARC_AER_MAXSPEC];  /* class of each species */
  const float *mass[ARC_AER_MAXBIN][ARC_AER_MAXSPEC];   /* chem(i,k,j,l): ug/kg-dryair */
  const float *num[ARC_AER_MAXBIN];          /* #/kg-dryair */
  float sigmag[ARC_AER_MAXBIN];              /* modal: geometric standard deviation of each mode (1.7, 2.0, 2.5) */
  const float *alt, *dz8w;                   /* inverse density m3/kg, layer thickness m (i,k,j) */
} ArcAerIn;

typedef struct ArcAerOut {
  float *tauaer[4], *gaer[4], *waer[4];      /* 300, 400, 600, 999 nm */
  float *tauaerlw[16];
  float *extaerlw[16];                       /* optional, 1/km */
} ArcAerOut;

/* refr / refi: [ARC_CLS_N][20] species refractive indices n + i k at the 4 SW wavelengths then the 16 LW band centres
 * (module_data_rrtmgaeropt.F in WRF-Chem); NULL selects the built-in representative values. Builds the Chebyshev-Mie tables. */
int  arc_aer_init(const float *refr, const float *refi);
/* copies the built-in refractive indices ([ARC_CLS_N][20] each); host only, no GPU needed */
void arc_aer_default_refindex(float *refr, float *refi);
int  arc_aer_optics(const ArcDims *d, const ArcAerIn *in, ArcAerOut *out);
/* test taps (host): Chebyshev-interpolated Q_ext, Q_sca, g of one sphere through the same tables (CPU evaluation of the
 * uploaded coefficients) and direct Mie theory for comparison */
int  arc_aer_table_eval(int wl, float radius_cm, float refr, float refi, float *qext, float *qsca, float *g);
int  arc_aer_mie_direct(int wl, float radius_cm, float refr, float refi, float *qext, float *qsca, float *g);

/* Kernel launch counter (number of this library's CUDA kernels launched since init) */
long long arc_rad_launch_count(void);
/* The library's main CUDA stream (cudaStream_t as void*): for event timing by the caller and for ordering the library behind
 * a producer (cudaStreamWaitEvent(arc_rad_stream(), ev) before a call, see "Streams" at the top).  Every call also uses two
 * more private streams that fork from and join this one inside the call. */
void *arc_rad_stream(void);
/* time (ms, CUDA events on the launching stream) spent in the named kernel class during the last call:
 * "sw_mcica", "sw_prep", "sw_solve", "sw_sweep", "sw_reduce", "lw_mcica", "lw_prep", "lw_solve", "lw_sweep", "lw_reduce",
 * "aer_optics"; returns <0 if unknown.  With the sweep overlap on, a class's time includes the slow-down from the kernels
 * of the other stream running beside it. */
float arc_rad_last_kernel_ms(const char *name);
/* Sweep overlap (default on, or $ARC_RAD_OVERLAP=0): the memory-bound sweep + reduce kernels of inner chunk k run on a second,
 * higher-priority stream while the compute-bound solver of chunk k+1 runs on the main stream.  Results are bit-identical
 * either way; off is for per-kernel timing.  Returns the previous setting. */
int arc_rad_set_overlap(int on);

/* Host-side layout logic exposed for the CPU tests (no CUDA device needed):
 * arc_rad_test_sweep_groups: the sweep groups (consecutive g-points of one band, at most gmax each) for bands with ng[b] g-points;
 *   writes band / first g-point (0-based, cumulative over the bands) / size of every group, returns the number of groups (<= 32).
 * arc_rad_test_coef_index: word index of (field, layer, column) in the tiled coefficient workspace of column capacity cap. */
int arc_rad_test_sweep_groups(const int *ng, int nbands, int gmax, int *band, int *g0, int *size);
long long arc_rad_test_coef_index(int field, int layer, long long column, long long cap, int nfields);

#ifdef __cplusplus
}
#endif
#endif /* ARC_RAD_H */
