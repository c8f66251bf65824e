/*
 * disort_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the DISORT 2.x / cdisort 2.1.3 discrete-ordinate solve that
 * pyharp reaches through disort::Disort::forward (reference call sites:
 * tests/test_disort.cpp:49, examples/amars_sw.cpp:280, examples/amars_lw.cpp:80,
 * src/radiation/radiation_band.cpp:124-127; legacy direct c_disort call
 * src/rtsolver/rt_solver_disort.cpp_:147).
 *
 * The arithmetic itself lives in an un-vendored third-party dependency,
 * github.com/zoeyzyhu/pydisort @ afee3ec897f (cmake/pydisort.cmake:9-11), which
 * wraps cdisort 2.1.3; neither is present under /root/reference.  This file
 * restates the published DISORT v2 algorithm (Stamnes, Tsay, Wiscombe, Jayaweera
 * 1988; Stamnes, Tsay, Wiscombe, Laszlo 2000 "DISORT report v1.1") routine by
 * routine.  PARITY AGAINST cdisort ITSELF IS UNPINNED; the oracle is pinned
 * against the DISORT test-problem golden vectors in tests/golden/ (SURVEY.md
 * Appendix B) and an independent numpy solve (tests/independent_solver.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product path never does.
 */
#ifndef DISORT_ORACLE_H
#define DISORT_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* flag bits (same values as include/b200disort.h) */
#define ORC_USRTAU 1
#define ORC_USRANG 2
#define ORC_LAMBER 4
#define ORC_PLANCK 8
#define ORC_ONLYFL 16
#define ORC_INTCOR 32     /* intensity_correction */
#define ORC_OLDINTCOR 64  /* old_intensity_correction (Nakajima-Tanaka) */

typedef struct {
  int nlyr, nstr, nmom; /* pmom has (nmom+1) entries per layer, k = 0..nmom */
  int ntau, numu, nphi; /* user levels / angles (ignored unless USRTAU/USRANG) */
  int flags;
  double accur;
  double wvnmlo, wvnmhi;
  /* boundary conditions (ds.bc.*) */
  double fbeam, umu0, phi0, fisot, albedo, btemp, ttemp, temis;
} orc_problem;

/* One c_disort-equivalent solve, DISORT ordering (layer 0 = top).
 * dtauc[nlyr], ssalb[nlyr], pmom[nlyr*(nmom+1)], temper[nlyr+1] (or NULL),
 * utau[ntau] / umu[numu] / phi[nphi] (degrees) user grids (may be NULL when unused).
 * Outputs, sized by the effective ntau (= nlyr+1 unless USRTAU):
 *   rfldir, rfldn, flup, dfdt, uavg [ntau]; uu[nphi][ntau][numu] (may be NULL if ONLYFL).
 * Returns 0 on success, >0 = status code (see b200disort.h). */
int orc_disort(const orc_problem* pb, const double* dtauc, const double* ssalb,
               const double* pmom, const double* temper, const double* utau,
               const double* umu, const double* phi, double* rfldir,
               double* rfldn, double* flup, double* dfdt, double* uavg,
               double* uu);

/* pydisort-style batched forward:
 * prop (nwave,ncol,nlyr,nprop) with index 0 = SURFACE layer; bc arrays are
 * (nwave,ncol) or NULL (=> 0; umu0 => 1, temis => 0); temf (ncol,nlyr+1)
 * index 0 = surface level; wave_lower/upper [nwave] or NULL.
 * flx (nwave,ncol,nlyr+1,2) [..,0]=flup, [..,1]=rfldir+rfldn, level 0 = surface.
 * uu  (nwave,ncol,nphi,ntau,numu) or NULL.  status (nwave,ncol) int.
 * nthreads<=1: serial loop exactly like the reference; >1: OpenMP over solves. */
typedef struct {
  const double *umu0, *fbeam, *albedo, *btemp, *ttemp, *temis, *fisot, *phi0;
} orc_bc;

int orc_forward(int nwave, int ncol, int nlyr, int nstr, int nmom, int nprop,
                int flags, double accur, int ntau, int numu, int nphi,
                const double* utau, const double* umu, const double* phi,
                const double* wave_lower, const double* wave_upper,
                const double* prop, const orc_bc* bc, const double* temf,
                double* flx, double* uu, int* status, int nthreads);

/* helpers exposed for unit tests */
void orc_qgausn(int m, double* gmu, double* gwt);
double orc_plkavg(double wnumlo, double wnumhi, double t);
int orc_asymtx(const double* aa, int n, double* evec, double* eval);
void orc_lepoly(int nmu, int m, int twonm1, const double* mu, double* ylm /* [(twonm1+1)][nmu] */);
int orc_real_bytes(void); /* sizeof(REAL) of the internal arithmetic (8 or 16) */

#ifdef __cplusplus
}
#endif
#endif
