/* Stub netcdf.h for building the reference sources without netcdf-c (TEST INFRASTRUCTURE).
 * Only the ".nc" branches of wghmStateFile.cpp / snowInElevationFile.cpp /
 * additionalOutputInputFile.cpp reference these symbols; the harness only ever uses the
 * ".txt" branches, so every entry point simply reports failure. */
#ifndef WGK_STUB_NETCDF_H
#define WGK_STUB_NETCDF_H
#include <stddef.h>
#define NC_NOWRITE 0
#define NC_NETCDF4 0x1000
#define NC_DOUBLE 6
#define NC_INT 4
#define NC_SHUFFLE 1
#define NC_NOERR 0
static inline int nc_open(const char *, int, int *) { return -1; }
static inline int nc_create(const char *, int, int *) { return -1; }
static inline int nc_close(int) { return -1; }
static inline int nc_inq_varid(int, const char *, int *) { return -1; }
static inline int nc_inq_vardimid(int, int, int *) { return -1; }
static inline int nc_inq_dimlen(int, int, size_t *) { return -1; }
static inline int nc_get_var_int(int, int, int *) { return -1; }
static inline int nc_get_var_double(int, int, double *) { return -1; }
static inline int nc_get_var1_double(int, int, const size_t *, double *) { return -1; }
static inline int nc_put_var_int(int, int, const int *) { return -1; }
static inline int nc_put_var_double(int, int, const double *) { return -1; }
static inline int nc_put_var1_double(int, int, const size_t *, const double *) { return -1; }
static inline int nc_def_dim(int, const char *, size_t, int *) { return -1; }
static inline int nc_def_var(int, const char *, int, int, const int *, int *) { return -1; }
static inline int nc_def_var_chunking(int, int, int, const size_t *) { return -1; }
static inline int nc_def_var_deflate(int, int, int, int, int) { return -1; }
static inline int nc_enddef(int) { return -1; }
static inline const char *nc_strerror(int) { return "netcdf stub: not available"; }
#endif
