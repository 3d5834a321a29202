/* hdf5_abi/hdf5.h — the slice of the public HDF5 ABI the filter plugin needs, for images
 * without libhdf5 development headers (this one).  When real HDF5 headers are available,
 * build with -DDRICE_USE_SYSTEM_HDF5 and this file is not used.  Declarations follow the
 * HDF5 public headers H5Zpublic.h / H5public.h (1.10+: hid_t is int64_t). */
#ifndef DRICE_HDF5_ABI_H
#define DRICE_HDF5_ABI_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef int     herr_t;
typedef int     htri_t;
typedef int64_t hid_t;
typedef int     H5Z_filter_t;

#define H5Z_CLASS_T_VERS   1
#define H5Z_FLAG_REVERSE   0x0100
#define H5Z_FLAG_OPTIONAL  0x0001

typedef htri_t (*H5Z_can_apply_func_t)(hid_t dcpl_id, hid_t type_id, hid_t space_id);
typedef herr_t (*H5Z_set_local_func_t)(hid_t dcpl_id, hid_t type_id, hid_t space_id);
typedef size_t (*H5Z_func_t)(unsigned int flags, size_t cd_nelmts, const unsigned int cd_values[],
                             size_t nbytes, size_t *buf_size, void **buf);

typedef struct H5Z_class2_t {
    int                  version;
    H5Z_filter_t         id;
    unsigned             encoder_present;
    unsigned             decoder_present;
    const char          *name;
    H5Z_can_apply_func_t can_apply;
    H5Z_set_local_func_t set_local;
    H5Z_func_t           filter;
} H5Z_class2_t;
#define H5Z_class_t H5Z_class2_t
#ifdef __cplusplus
}
#endif
#endif
