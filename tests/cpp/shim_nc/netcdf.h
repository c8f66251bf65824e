/* Minimal stand-in for <netcdf.h> -- TEST INFRASTRUCTURE ONLY.  The NetCDF-C library is absent from this image
 * and the reference ships no .nc file, so the NetCDF branches of the reference's
 * src/opacity/rfm.cpp:34-120 and src/utils/read_weights.cpp:18-46 could not run.  This header implements exactly
 * the calls those two files make (nc_open, nc_inq_dimid, nc_inq_dimlen, nc_inq_varid, nc_get_var_double, nc_close,
 * nc_strerror) over a flat container file written by tests/nc_container.py, so that examples/amars_lw.cpp and the
 * RFM attenuator compile and run UNCHANGED from where they lie.  Plain C: it is included inside extern "C" {}.
 *
 * Container layout (little endian):  "B200NC01" | int32 ndims | ndims x { char name[64]; int64 len } |
 *   int32 nvars | nvars x { char name[64]; int64 nelem; int64 offset } | float64 payloads at their offsets. */
#ifndef B200_NETCDF_SHIM_H
#define B200_NETCDF_SHIM_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NC_NOERR 0
#define NC_NOWRITE 0
#define NC_NETCDF4 0x1000
#define NC_EBADID (-33)
#define NC_EBADDIM (-46)
#define NC_ENOTVAR (-49)
#define NC_ENOTNC (-51)

#define B200_NC_MAX_FILES 16
#define B200_NC_MAX_ITEMS 64

typedef struct {
  FILE* f;
  int ndims, nvars;
  char dname[B200_NC_MAX_ITEMS][64];
  long long dlen[B200_NC_MAX_ITEMS];
  char vname[B200_NC_MAX_ITEMS][64];
  long long vnelem[B200_NC_MAX_ITEMS], voff[B200_NC_MAX_ITEMS];
} b200_nc_file;

static b200_nc_file* b200_nc_slot(int ncid) {
  static b200_nc_file files[B200_NC_MAX_FILES];
  return (ncid >= 0 && ncid < B200_NC_MAX_FILES) ? &files[ncid] : NULL;
}

static int nc_open(const char* path, int mode, int* ncidp) {
  int id, i;
  char magic[8];
  int n = 0;
  (void)mode;
  for (id = 0; id < B200_NC_MAX_FILES; ++id)
    if (!b200_nc_slot(id)->f) break;
  if (id == B200_NC_MAX_FILES) return NC_EBADID;
  b200_nc_file* s = b200_nc_slot(id);
  s->f = fopen(path, "rb");
  if (!s->f) return NC_EBADID;
  if (fread(magic, 1, 8, s->f) != 8 || memcmp(magic, "B200NC01", 8) != 0) { fclose(s->f); s->f = NULL; return NC_ENOTNC; }
  if (fread(&n, 4, 1, s->f) != 1 || n < 0 || n > B200_NC_MAX_ITEMS) { fclose(s->f); s->f = NULL; return NC_ENOTNC; }
  s->ndims = n;
  for (i = 0; i < n; ++i)
    if (fread(s->dname[i], 1, 64, s->f) != 64 || fread(&s->dlen[i], 8, 1, s->f) != 1) { fclose(s->f); s->f = NULL; return NC_ENOTNC; }
  if (fread(&n, 4, 1, s->f) != 1 || n < 0 || n > B200_NC_MAX_ITEMS) { fclose(s->f); s->f = NULL; return NC_ENOTNC; }
  s->nvars = n;
  for (i = 0; i < n; ++i)
    if (fread(s->vname[i], 1, 64, s->f) != 64 || fread(&s->vnelem[i], 8, 1, s->f) != 1 || fread(&s->voff[i], 8, 1, s->f) != 1) {
      fclose(s->f); s->f = NULL; return NC_ENOTNC;
    }
  *ncidp = id;
  return NC_NOERR;
}

static int nc_close(int ncid) {
  b200_nc_file* s = b200_nc_slot(ncid);
  if (!s || !s->f) return NC_EBADID;
  fclose(s->f);
  s->f = NULL;
  return NC_NOERR;
}

static int nc_inq_dimid(int ncid, const char* name, int* idp) {
  b200_nc_file* s = b200_nc_slot(ncid);
  int i;
  if (!s || !s->f) return NC_EBADID;
  for (i = 0; i < s->ndims; ++i)
    if (strncmp(s->dname[i], name, 64) == 0) { *idp = i; return NC_NOERR; }
  return NC_EBADDIM;
}

static int nc_inq_dimlen(int ncid, int dimid, size_t* lenp) {
  b200_nc_file* s = b200_nc_slot(ncid);
  if (!s || !s->f) return NC_EBADID;
  if (dimid < 0 || dimid >= s->ndims) return NC_EBADDIM;
  *lenp = (size_t)s->dlen[dimid];
  return NC_NOERR;
}

static int nc_inq_varid(int ncid, const char* name, int* idp) {
  b200_nc_file* s = b200_nc_slot(ncid);
  int i;
  if (!s || !s->f) return NC_EBADID;
  for (i = 0; i < s->nvars; ++i)
    if (strncmp(s->vname[i], name, 64) == 0) { *idp = i; return NC_NOERR; }
  return NC_ENOTVAR;
}

static int nc_get_var_double(int ncid, int varid, double* out) {
  b200_nc_file* s = b200_nc_slot(ncid);
  if (!s || !s->f) return NC_EBADID;
  if (varid < 0 || varid >= s->nvars) return NC_ENOTVAR;
  if (fseek(s->f, (long)s->voff[varid], SEEK_SET) != 0) return NC_ENOTNC;
  if (fread(out, 8, (size_t)s->vnelem[varid], s->f) != (size_t)s->vnelem[varid]) return NC_ENOTNC;
  return NC_NOERR;
}

static const char* nc_strerror(int err) {
  switch (err) {
    case NC_NOERR: return "No error";
    case NC_EBADID: return "NetCDF: Not a valid ID (file missing?)";
    case NC_EBADDIM: return "NetCDF: Invalid dimension ID or name";
    case NC_ENOTVAR: return "NetCDF: Variable not found";
    case NC_ENOTNC: return "NetCDF: Unknown file format";
    default: return "NetCDF: error";
  }
}
#endif
