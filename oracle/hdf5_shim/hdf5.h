/* Minimal stand-in for <hdf5.h>: ONLY what the reference codec translation unit
 * touches (reference src/deltaRice.c:8,19-28,473,496), so the unmodified reference
 * source can be compiled without libhdf5 (absent from this image).
 * Test infrastructure only.  Layout of H5Z_class2_t follows the public HDF5 ABI. */
#ifndef DRICE_ORACLE_HDF5_SHIM_H
#define DRICE_ORACLE_HDF5_SHIM_H
#include <stddef.h>
#include <stdint.h>

typedef int     herr_t;
typedef int     htri_t;
typedef int64_t hid_t;
typedef int     H5Z_filter_t;

#define H5Z_CLASS_T_VERS 1
#define H5Z_FLAG_REVERSE 0x0100

typedef htri_t (*H5Z_can_apply_func_t)(hid_t dcpl, hid_t type, hid_t space);
typedef herr_t (*H5Z_set_local_func_t)(hid_t dcpl, hid_t type, hid_t space);
typedef size_t (*H5Z_func_t)(unsigned int flags, size_t cd_nelmts,
                             const unsigned int cd_values[], size_t nbytes,
                             size_t *buf_size, void **buf);

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

/* the oracle never talks to a real libhdf5 */
static inline herr_t H5Zregister(const void *cls) { (void)cls; return 0; }
#endif
