/* oracle/shim/netcdf.h -- TEST INFRASTRUCTURE ONLY.
 * Minimal stand-in for the netCDF C API so that the reference's own sources
 * (read_atmos_data.c, make_in_and_outfiles.c, close_files.c) compile in a
 * container without libnetcdf.  Every call fails; the oracle harness only
 * uses FORCE_FORMAT ASCII/BINARY so none of them is ever reached. */
#ifndef VIC_ORACLE_SHIM_NETCDF_H
#define VIC_ORACLE_SHIM_NETCDF_H
#include <stddef.h>
#define NC_NOERR 0
#define NC_NOWRITE 0
#define NC_MAX_NAME 256
#define NC_MAX_VAR_DIMS 1024
#define NC_BYTE 1
#define NC_CHAR 2
#define NC_SHORT 3
#define NC_INT 4
#define NC_FLOAT 5
#define NC_DOUBLE 6
#define NC_UBYTE 7
#define NC_USHORT 8
typedef int nc_type;
static inline const char *nc_strerror(int) { return "netCDF unavailable (oracle shim)"; }
static inline int nc_open(const char *, int, int *) { return -1; }
static inline int nc_close(int) { return -1; }
static inline int nc_inq_varid(int, const char *, int *) { return -1; }
static inline int nc_inq_varndims(int, int, int *) { return -1; }
static inline int nc_inq_vardimid(int, int, int *) { return -1; }
static inline int nc_inq_vartype(int, int, nc_type *) { return -1; }
static inline int nc_inq_dimlen(int, int, size_t *) { return -1; }
static inline int nc_inq_dim(int, int, char *, size_t *) { return -1; }
static inline int nc_inq_var(int, int, char *, nc_type *, int *, int *, int *) { return -1; }
static inline int nc_get_att_float(int, int, const char *, float *) { return -1; }
static inline int nc_get_vara_double(int, int, const size_t *, const size_t *, double *) { return -1; }
static inline int nc_get_varm_double(int, int, const size_t *, const size_t *, const ptrdiff_t *, const ptrdiff_t *, double *) { return -1; }
static inline int nc_get_varm_float(int, int, const size_t *, const size_t *, const ptrdiff_t *, const ptrdiff_t *, float *) { return -1; }
static inline int nc_get_varm_short(int, int, const size_t *, const size_t *, const ptrdiff_t *, const ptrdiff_t *, short *) { return -1; }
static inline int nc_get_varm_ushort(int, int, const size_t *, const size_t *, const ptrdiff_t *, const ptrdiff_t *, unsigned short *) { return -1; }
#endif
