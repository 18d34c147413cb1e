/*
 * oracle/shim/hdf5.h -- DECLARATION-ONLY stand-in for the HDF5 C API.
 *
 * TEST INFRASTRUCTURE ONLY (oracle build).  The reference's headers
 * (src/utils.h, src/hdf5_tools.hpp, src/signal_batch.cc) include <hdf5.h>
 * because the reference can also read FAST5 containers.  The mapping hot path
 * never touches HDF5, and every input this repo feeds the reference is BLOW5,
 * so instead of running the vendored HDF5's autotools build (configure +
 * generated H5pubconf.h) the oracle build compiles the reference sources
 * against these prototypes and links oracle/shim/hdf5_stub.c, whose functions
 * abort with a message if a FAST5 code path is ever reached.
 *
 * Only the ~110 names the reference's sources mention are declared; the
 * prototypes follow the public HDF5 1.10 API.
 */
#ifndef SIGMAP_ORACLE_HDF5_SHIM_H
#define SIGMAP_ORACLE_HDF5_SHIM_H

#include <stddef.h>
#include <stdint.h>
#include <sys/types.h>

#ifdef __cplusplus
extern "C" {
#endif

#define H5_VERS_MAJOR 1
#define H5_VERS_MINOR 10
#define H5_VERS_RELEASE 6
#define H5_VERSION_GE(Maj, Min, Rel)                                         \
  (((H5_VERS_MAJOR == Maj) && (H5_VERS_MINOR == Min) &&                      \
    (H5_VERS_RELEASE >= Rel)) ||                                             \
   ((H5_VERS_MAJOR == Maj) && (H5_VERS_MINOR > Min)) || (H5_VERS_MAJOR > Maj))

typedef int64_t hid_t;
typedef int herr_t;
typedef int htri_t;
typedef unsigned int hbool_t;
typedef unsigned long long hsize_t;
typedef signed long long hssize_t;
typedef uint64_t haddr_t;

/* ---- property lists / misc constants ---- */
#define H5P_DEFAULT ((hid_t)0)
extern hid_t H5P_CLS_LINK_CREATE_ID_g;
#define H5P_LINK_CREATE (H5P_CLS_LINK_CREATE_ID_g)
#define H5S_ALL ((hid_t)0)

#define H5F_ACC_RDONLY (0x0000u)
#define H5F_ACC_RDWR (0x0001u)
#define H5F_ACC_TRUNC (0x0002u)
#define H5F_ACC_EXCL (0x0004u)
#define H5F_OBJ_FILE (0x0001u)
#define H5F_OBJ_DATASET (0x0002u)
#define H5F_OBJ_GROUP (0x0004u)
#define H5F_OBJ_DATATYPE (0x0008u)
#define H5F_OBJ_ATTR (0x0010u)
#define H5F_OBJ_ALL \
  (H5F_OBJ_FILE | H5F_OBJ_DATASET | H5F_OBJ_GROUP | H5F_OBJ_DATATYPE | H5F_OBJ_ATTR)
#define H5F_OBJ_LOCAL (0x0020u)

typedef enum H5_index_t {
  H5_INDEX_UNKNOWN = -1,
  H5_INDEX_NAME,
  H5_INDEX_CRT_ORDER,
  H5_INDEX_N
} H5_index_t;
typedef enum H5_iter_order_t {
  H5_ITER_UNKNOWN = -1,
  H5_ITER_INC,
  H5_ITER_DEC,
  H5_ITER_NATIVE,
  H5_ITER_N
} H5_iter_order_t;

/* ---- dataspaces ---- */
typedef enum H5S_class_t {
  H5S_NO_CLASS = -1,
  H5S_SCALAR = 0,
  H5S_SIMPLE = 1,
  H5S_NULL = 2
} H5S_class_t;

/* ---- datatypes ---- */
typedef enum H5T_class_t {
  H5T_NO_CLASS = -1,
  H5T_INTEGER = 0,
  H5T_FLOAT = 1,
  H5T_TIME = 2,
  H5T_STRING = 3,
  H5T_BITFIELD = 4,
  H5T_OPAQUE = 5,
  H5T_COMPOUND = 6,
  H5T_REFERENCE = 7,
  H5T_ENUM = 8,
  H5T_VLEN = 9,
  H5T_ARRAY = 10,
  H5T_NCLASSES
} H5T_class_t;
typedef enum H5T_sign_t {
  H5T_SGN_ERROR = -1,
  H5T_SGN_NONE = 0,
  H5T_SGN_2 = 1,
  H5T_NSGN = 2
} H5T_sign_t;
typedef enum H5T_cset_t {
  H5T_CSET_ERROR = -1,
  H5T_CSET_ASCII = 0,
  H5T_CSET_UTF8 = 1
} H5T_cset_t;
typedef enum H5T_direction_t {
  H5T_DIR_DEFAULT = 0,
  H5T_DIR_ASCEND = 1,
  H5T_DIR_DESCEND = 2
} H5T_direction_t;
#define H5T_VARIABLE ((size_t)(-1))

extern hid_t H5T_C_S1_g;
extern hid_t H5T_NATIVE_SCHAR_g, H5T_NATIVE_UCHAR_g, H5T_NATIVE_SHORT_g,
    H5T_NATIVE_USHORT_g, H5T_NATIVE_INT_g, H5T_NATIVE_UINT_g,
    H5T_NATIVE_LONG_g, H5T_NATIVE_ULONG_g, H5T_NATIVE_LLONG_g,
    H5T_NATIVE_ULLONG_g, H5T_NATIVE_FLOAT_g, H5T_NATIVE_DOUBLE_g,
    H5T_NATIVE_LDOUBLE_g;
#define H5T_C_S1 (H5T_C_S1_g)
#define H5T_NATIVE_CHAR (H5T_NATIVE_SCHAR_g)
#define H5T_NATIVE_UCHAR (H5T_NATIVE_UCHAR_g)
#define H5T_NATIVE_SHORT (H5T_NATIVE_SHORT_g)
#define H5T_NATIVE_USHORT (H5T_NATIVE_USHORT_g)
#define H5T_NATIVE_INT (H5T_NATIVE_INT_g)
#define H5T_NATIVE_UINT (H5T_NATIVE_UINT_g)
#define H5T_NATIVE_LONG (H5T_NATIVE_LONG_g)
#define H5T_NATIVE_ULONG (H5T_NATIVE_ULONG_g)
#define H5T_NATIVE_LLONG (H5T_NATIVE_LLONG_g)
#define H5T_NATIVE_ULLONG (H5T_NATIVE_ULLONG_g)
#define H5T_NATIVE_FLOAT (H5T_NATIVE_FLOAT_g)
#define H5T_NATIVE_DOUBLE (H5T_NATIVE_DOUBLE_g)
#define H5T_NATIVE_LDOUBLE (H5T_NATIVE_LDOUBLE_g)

/* ---- groups / objects ---- */
typedef struct H5G_info_t {
  int storage_type;
  hsize_t nlinks;
  int64_t max_corder;
  hbool_t mounted;
} H5G_info_t;
typedef enum H5O_type_t {
  H5O_TYPE_UNKNOWN = -1,
  H5O_TYPE_GROUP,
  H5O_TYPE_DATASET,
  H5O_TYPE_NAMED_DATATYPE,
  H5O_TYPE_NTYPES
} H5O_type_t;
typedef struct H5O_info_t {
  unsigned long fileno;
  haddr_t addr;
  H5O_type_t type;
  unsigned rc;
  hsize_t num_attrs;
} H5O_info_t;
#define H5O_INFO_BASIC 0x0001u
#define H5O_INFO_NUM_ATTRS 0x0004u

/* ---- H5A ---- */
herr_t H5Aclose(hid_t attr_id);
hid_t H5Acreate2(hid_t loc_id, const char *attr_name, hid_t type_id,
                 hid_t space_id, hid_t acpl_id, hid_t aapl_id);
htri_t H5Aexists_by_name(hid_t obj_id, const char *obj_name,
                         const char *attr_name, hid_t lapl_id);
ssize_t H5Aget_name_by_idx(hid_t loc_id, const char *obj_name,
                           H5_index_t idx_type, H5_iter_order_t order,
                           hsize_t n, char *name, size_t size, hid_t lapl_id);
hid_t H5Aget_space(hid_t attr_id);
hid_t H5Aget_type(hid_t attr_id);
hsize_t H5Aget_storage_size(hid_t attr_id);
hid_t H5Aopen(hid_t obj_id, const char *attr_name, hid_t aapl_id);
hid_t H5Aopen_by_name(hid_t loc_id, const char *obj_name,
                      const char *attr_name, hid_t aapl_id, hid_t lapl_id);
herr_t H5Aread(hid_t attr_id, hid_t type_id, void *buf);
herr_t H5Awrite(hid_t attr_id, hid_t type_id, const void *buf);
/* ---- H5D ---- */
herr_t H5Dclose(hid_t dset_id);
hid_t H5Dcreate2(hid_t loc_id, const char *name, hid_t type_id, hid_t space_id,
                 hid_t lcpl_id, hid_t dcpl_id, hid_t dapl_id);
hid_t H5Dget_space(hid_t dset_id);
hid_t H5Dget_type(hid_t dset_id);
hid_t H5Dopen2(hid_t file_id, const char *name, hid_t dapl_id);
#define H5Dopen H5Dopen2
herr_t H5Dread(hid_t dset_id, hid_t mem_type_id, hid_t mem_space_id,
               hid_t file_space_id, hid_t plist_id, void *buf);
herr_t H5Dvlen_reclaim(hid_t type_id, hid_t space_id, hid_t plist_id,
                       void *buf);
herr_t H5Dwrite(hid_t dset_id, hid_t mem_type_id, hid_t mem_space_id,
                hid_t file_space_id, hid_t plist_id, const void *buf);
/* ---- H5F ---- */
herr_t H5Fclose(hid_t file_id);
hid_t H5Fcreate(const char *filename, unsigned flags, hid_t create_plist,
                hid_t access_plist);
ssize_t H5Fget_obj_count(hid_t file_id, unsigned types);
htri_t H5Fis_hdf5(const char *filename);
hid_t H5Fopen(const char *filename, unsigned flags, hid_t access_plist);
/* ---- H5G ---- */
herr_t H5Gclose(hid_t group_id);
hid_t H5Gcreate2(hid_t loc_id, const char *name, hid_t lcpl_id, hid_t gcpl_id,
                 hid_t gapl_id);
herr_t H5Gget_info(hid_t loc_id, H5G_info_t *ginfo);
hid_t H5Gopen2(hid_t loc_id, const char *name, hid_t gapl_id);
#define H5Gopen H5Gopen2
/* ---- H5L ---- */
htri_t H5Lexists(hid_t loc_id, const char *name, hid_t lapl_id);
ssize_t H5Lget_name_by_idx(hid_t loc_id, const char *group_name,
                           H5_index_t idx_type, H5_iter_order_t order,
                           hsize_t n, char *name, size_t size, hid_t lapl_id);
/* ---- H5O ---- */
herr_t H5Oclose(hid_t object_id);
htri_t H5Oexists_by_name(hid_t loc_id, const char *name, hid_t lapl_id);
herr_t H5Oget_info(hid_t loc_id, H5O_info_t *oinfo);
hid_t H5Oopen(hid_t loc_id, const char *name, hid_t lapl_id);
/* ---- H5P ---- */
herr_t H5Pclose(hid_t plist_id);
hid_t H5Pcreate(hid_t cls_id);
herr_t H5Pset_create_intermediate_group(hid_t plist_id, unsigned crt_intmd);
/* ---- H5S ---- */
herr_t H5Sclose(hid_t space_id);
hid_t H5Screate(H5S_class_t type);
hid_t H5Screate_simple(int rank, const hsize_t dims[], const hsize_t maxdims[]);
int H5Sget_simple_extent_dims(hid_t space_id, hsize_t dims[],
                              hsize_t maxdims[]);
int H5Sget_simple_extent_ndims(hid_t space_id);
H5S_class_t H5Sget_simple_extent_type(hid_t space_id);
/* ---- H5T ---- */
herr_t H5Tclose(hid_t type_id);
hid_t H5Tcopy(hid_t type_id);
hid_t H5Tcreate(H5T_class_t type, size_t size);
H5T_class_t H5Tget_class(hid_t type_id);
H5T_cset_t H5Tget_cset(hid_t type_id);
int H5Tget_member_index(hid_t type_id, const char *name);
char *H5Tget_member_name(hid_t type_id, unsigned membno);
hid_t H5Tget_member_type(hid_t type_id, unsigned membno);
hid_t H5Tget_native_type(hid_t type_id, H5T_direction_t direction);
int H5Tget_nmembers(hid_t type_id);
H5T_sign_t H5Tget_sign(hid_t type_id);
size_t H5Tget_size(hid_t type_id);
herr_t H5Tinsert(hid_t parent_id, const char *name, size_t offset,
                 hid_t member_id);
htri_t H5Tis_variable_str(hid_t type_id);
herr_t H5Tset_cset(hid_t type_id, H5T_cset_t cset);
herr_t H5Tset_size(hid_t type_id, size_t size);

#ifdef __cplusplus
}
#endif
#endif /* SIGMAP_ORACLE_HDF5_SHIM_H */
