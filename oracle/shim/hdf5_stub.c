/*
 * oracle/shim/hdf5_stub.c -- link-time stand-ins for the HDF5 functions that
 * oracle/shim/hdf5.h declares.  TEST INFRASTRUCTURE ONLY.  The oracle build of
 * the reference reads BLOW5 exclusively; reaching any of these means a FAST5
 * path was taken, which this build does not support, so they abort loudly.
 * (Generated once from hdf5.h's prototype list; regenerate if that changes.)
 */
#include <stdio.h>
#include <stdlib.h>
#include "hdf5.h"

static void sigmap_oracle_no_hdf5(const char *fn) {
  fprintf(stderr, "oracle build: HDF5/FAST5 is not supported (called %s); use BLOW5 input\n", fn);
  abort();
}

hid_t H5P_CLS_LINK_CREATE_ID_g = -1;
hid_t H5T_C_S1_g = -1;
hid_t H5T_NATIVE_SCHAR_g = -1, H5T_NATIVE_UCHAR_g = -1, H5T_NATIVE_SHORT_g = -1,
      H5T_NATIVE_USHORT_g = -1, H5T_NATIVE_INT_g = -1, H5T_NATIVE_UINT_g = -1,
      H5T_NATIVE_LONG_g = -1, H5T_NATIVE_ULONG_g = -1, H5T_NATIVE_LLONG_g = -1,
      H5T_NATIVE_ULLONG_g = -1, H5T_NATIVE_FLOAT_g = -1, H5T_NATIVE_DOUBLE_g = -1,
      H5T_NATIVE_LDOUBLE_g = -1;

herr_t H5Aclose(hid_t attr_id) { sigmap_oracle_no_hdf5("H5Aclose"); return (herr_t)0; }
hid_t H5Acreate2(hid_t loc_id, const char *attr_name, hid_t type_id, hid_t space_id, hid_t acpl_id, hid_t aapl_id) { sigmap_oracle_no_hdf5("H5Acreate2"); return (hid_t)0; }
htri_t H5Aexists_by_name(hid_t obj_id, const char *obj_name, const char *attr_name, hid_t lapl_id) { sigmap_oracle_no_hdf5("H5Aexists_by_name"); return (htri_t)0; }
ssize_t H5Aget_name_by_idx(hid_t loc_id, const char *obj_name, H5_index_t idx_type, H5_iter_order_t order, hsize_t n, char *name, size_t size, hid_t lapl_id) { sigmap_oracle_no_hdf5("H5Aget_name_by_idx"); return (ssize_t)0; }
hid_t H5Aget_space(hid_t attr_id) { sigmap_oracle_no_hdf5("H5Aget_space"); return (hid_t)0; }
hid_t H5Aget_type(hid_t attr_id) { sigmap_oracle_no_hdf5("H5Aget_type"); return (hid_t)0; }
hsize_t H5Aget_storage_size(hid_t attr_id) { sigmap_oracle_no_hdf5("H5Aget_storage_size"); return (hsize_t)0; }
hid_t H5Aopen(hid_t obj_id, const char *attr_name, hid_t aapl_id) { sigmap_oracle_no_hdf5("H5Aopen"); return (hid_t)0; }
hid_t H5Aopen_by_name(hid_t loc_id, const char *obj_name, const char *attr_name, hid_t aapl_id, hid_t lapl_id) { sigmap_oracle_no_hdf5("H5Aopen_by_name"); return (hid_t)0; }
herr_t H5Aread(hid_t attr_id, hid_t type_id, void *buf) { sigmap_oracle_no_hdf5("H5Aread"); return (herr_t)0; }
herr_t H5Awrite(hid_t attr_id, hid_t type_id, const void *buf) { sigmap_oracle_no_hdf5("H5Awrite"); return (herr_t)0; }
herr_t H5Dclose(hid_t dset_id) { sigmap_oracle_no_hdf5("H5Dclose"); return (herr_t)0; }
hid_t H5Dcreate2(hid_t loc_id, const char *name, hid_t type_id, hid_t space_id, hid_t lcpl_id, hid_t dcpl_id, hid_t dapl_id) { sigmap_oracle_no_hdf5("H5Dcreate2"); return (hid_t)0; }
hid_t H5Dget_space(hid_t dset_id) { sigmap_oracle_no_hdf5("H5Dget_space"); return (hid_t)0; }
hid_t H5Dget_type(hid_t dset_id) { sigmap_oracle_no_hdf5("H5Dget_type"); return (hid_t)0; }
hid_t H5Dopen2(hid_t file_id, const char *name, hid_t dapl_id) { sigmap_oracle_no_hdf5("H5Dopen2"); return (hid_t)0; }
herr_t H5Dread(hid_t dset_id, hid_t mem_type_id, hid_t mem_space_id, hid_t file_space_id, hid_t plist_id, void *buf) { sigmap_oracle_no_hdf5("H5Dread"); return (herr_t)0; }
herr_t H5Dvlen_reclaim(hid_t type_id, hid_t space_id, hid_t plist_id, void *buf) { sigmap_oracle_no_hdf5("H5Dvlen_reclaim"); return (herr_t)0; }
herr_t H5Dwrite(hid_t dset_id, hid_t mem_type_id, hid_t mem_space_id, hid_t file_space_id, hid_t plist_id, const void *buf) { sigmap_oracle_no_hdf5("H5Dwrite"); return (herr_t)0; }
herr_t H5Fclose(hid_t file_id) { sigmap_oracle_no_hdf5("H5Fclose"); return (herr_t)0; }
hid_t H5Fcreate(const char *filename, unsigned flags, hid_t create_plist, hid_t access_plist) { sigmap_oracle_no_hdf5("H5Fcreate"); return (hid_t)0; }
ssize_t H5Fget_obj_count(hid_t file_id, unsigned types) { sigmap_oracle_no_hdf5("H5Fget_obj_count"); return (ssize_t)0; }
htri_t H5Fis_hdf5(const char *filename) { sigmap_oracle_no_hdf5("H5Fis_hdf5"); return (htri_t)0; }
hid_t H5Fopen(const char *filename, unsigned flags, hid_t access_plist) { sigmap_oracle_no_hdf5("H5Fopen"); return (hid_t)0; }
herr_t H5Gclose(hid_t group_id) { sigmap_oracle_no_hdf5("H5Gclose"); return (herr_t)0; }
hid_t H5Gcreate2(hid_t loc_id, const char *name, hid_t lcpl_id, hid_t gcpl_id, hid_t gapl_id) { sigmap_oracle_no_hdf5("H5Gcreate2"); return (hid_t)0; }
herr_t H5Gget_info(hid_t loc_id, H5G_info_t *ginfo) { sigmap_oracle_no_hdf5("H5Gget_info"); return (herr_t)0; }
hid_t H5Gopen2(hid_t loc_id, const char *name, hid_t gapl_id) { sigmap_oracle_no_hdf5("H5Gopen2"); return (hid_t)0; }
htri_t H5Lexists(hid_t loc_id, const char *name, hid_t lapl_id) { sigmap_oracle_no_hdf5("H5Lexists"); return (htri_t)0; }
ssize_t H5Lget_name_by_idx(hid_t loc_id, const char *group_name, H5_index_t idx_type, H5_iter_order_t order, hsize_t n, char *name, size_t size, hid_t lapl_id) { sigmap_oracle_no_hdf5("H5Lget_name_by_idx"); return (ssize_t)0; }
herr_t H5Oclose(hid_t object_id) { sigmap_oracle_no_hdf5("H5Oclose"); return (herr_t)0; }
htri_t H5Oexists_by_name(hid_t loc_id, const char *name, hid_t lapl_id) { sigmap_oracle_no_hdf5("H5Oexists_by_name"); return (htri_t)0; }
herr_t H5Oget_info(hid_t loc_id, H5O_info_t *oinfo) { sigmap_oracle_no_hdf5("H5Oget_info"); return (herr_t)0; }
hid_t H5Oopen(hid_t loc_id, const char *name, hid_t lapl_id) { sigmap_oracle_no_hdf5("H5Oopen"); return (hid_t)0; }
herr_t H5Pclose(hid_t plist_id) { sigmap_oracle_no_hdf5("H5Pclose"); return (herr_t)0; }
hid_t H5Pcreate(hid_t cls_id) { sigmap_oracle_no_hdf5("H5Pcreate"); return (hid_t)0; }
herr_t H5Pset_create_intermediate_group(hid_t plist_id, unsigned crt_intmd) { sigmap_oracle_no_hdf5("H5Pset_create_intermediate_group"); return (herr_t)0; }
herr_t H5Sclose(hid_t space_id) { sigmap_oracle_no_hdf5("H5Sclose"); return (herr_t)0; }
hid_t H5Screate(H5S_class_t type) { sigmap_oracle_no_hdf5("H5Screate"); return (hid_t)0; }
hid_t H5Screate_simple(int rank, const hsize_t dims[], const hsize_t maxdims[]) { sigmap_oracle_no_hdf5("H5Screate_simple"); return (hid_t)0; }
int H5Sget_simple_extent_dims(hid_t space_id, hsize_t dims[], hsize_t maxdims[]) { sigmap_oracle_no_hdf5("H5Sget_simple_extent_dims"); return (int)0; }
int H5Sget_simple_extent_ndims(hid_t space_id) { sigmap_oracle_no_hdf5("H5Sget_simple_extent_ndims"); return (int)0; }
H5S_class_t H5Sget_simple_extent_type(hid_t space_id) { sigmap_oracle_no_hdf5("H5Sget_simple_extent_type"); return (H5S_class_t)0; }
herr_t H5Tclose(hid_t type_id) { sigmap_oracle_no_hdf5("H5Tclose"); return (herr_t)0; }
hid_t H5Tcopy(hid_t type_id) { sigmap_oracle_no_hdf5("H5Tcopy"); return (hid_t)0; }
hid_t H5Tcreate(H5T_class_t type, size_t size) { sigmap_oracle_no_hdf5("H5Tcreate"); return (hid_t)0; }
H5T_class_t H5Tget_class(hid_t type_id) { sigmap_oracle_no_hdf5("H5Tget_class"); return (H5T_class_t)0; }
H5T_cset_t H5Tget_cset(hid_t type_id) { sigmap_oracle_no_hdf5("H5Tget_cset"); return (H5T_cset_t)0; }
int H5Tget_member_index(hid_t type_id, const char *name) { sigmap_oracle_no_hdf5("H5Tget_member_index"); return (int)0; }
char * H5Tget_member_name(hid_t type_id, unsigned membno) { sigmap_oracle_no_hdf5("H5Tget_member_name"); return (char *)0; }
hid_t H5Tget_member_type(hid_t type_id, unsigned membno) { sigmap_oracle_no_hdf5("H5Tget_member_type"); return (hid_t)0; }
hid_t H5Tget_native_type(hid_t type_id, H5T_direction_t direction) { sigmap_oracle_no_hdf5("H5Tget_native_type"); return (hid_t)0; }
int H5Tget_nmembers(hid_t type_id) { sigmap_oracle_no_hdf5("H5Tget_nmembers"); return (int)0; }
H5T_sign_t H5Tget_sign(hid_t type_id) { sigmap_oracle_no_hdf5("H5Tget_sign"); return (H5T_sign_t)0; }
size_t H5Tget_size(hid_t type_id) { sigmap_oracle_no_hdf5("H5Tget_size"); return (size_t)0; }
herr_t H5Tinsert(hid_t parent_id, const char *name, size_t offset, hid_t member_id) { sigmap_oracle_no_hdf5("H5Tinsert"); return (herr_t)0; }
htri_t H5Tis_variable_str(hid_t type_id) { sigmap_oracle_no_hdf5("H5Tis_variable_str"); return (htri_t)0; }
herr_t H5Tset_cset(hid_t type_id, H5T_cset_t cset) { sigmap_oracle_no_hdf5("H5Tset_cset"); return (herr_t)0; }
herr_t H5Tset_size(hid_t type_id, size_t size) { sigmap_oracle_no_hdf5("H5Tset_size"); return (herr_t)0; }
