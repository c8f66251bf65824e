/*
 * cdisort213/cdisort.h -- the part of cdisort 2.1.3's public header that pyharp touches
 * (included by src/rtsolver/rtsolver.hpp:16; fields used at tests/test_disort.cpp:21-23,35,
 * examples/amars_sw.cpp:56-62, src/radiation/radiation_band.cpp:58,
 * src/rtsolver/rt_solver_disort.cpp_:67-95,139-140,176-181,282 and
 * src/radiation/radiation_band_exchanger.cpp_:35,71-94).
 *
 * In this project the structs are descriptors only: they carry the problem shape, flags and
 * user grids across the disort:: C++ surface.  There is no c_disort() behind them -- every
 * solve runs in the CUDA library through include/b200disort.h.
 */
#ifndef B200_CDISORT_H
#define B200_CDISORT_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int ibcnd, usrtau, usrang, lamber, planck, spher, onlyfl;
  int prnt[5];
  int brdf_type, quiet, intensity_correction, old_intensity_correction, general_source, output_uum;
} disort_flag;

typedef struct {
  double fbeam, umu0, phi0, fisot, fluor, albedo, btemp, ttemp, temis;
} disort_bc;

typedef struct {
  double rfldir, rfldn, flup, dfdt, uavg, uavgdn, uavgup, uavgso;
} disort_radiant;

typedef struct {
  char header[128];
  disort_flag flag;
  disort_bc bc;
  int nlyr, nmom, nstr, nmom_nstr, ntau, numu, nphi;
  double wvnmlo, wvnmhi, accur, radius;
  double *dtauc, *ssalb, *pmom, *temper, *utau, *umu, *phi, *zd;
} disort_state;

typedef struct {
  disort_radiant* rad;
  double *albmed, *trnmed, *uu, *u0u, *uum;
} disort_output;

#ifdef __cplusplus
}
#endif
#endif /* B200_CDISORT_H */
