/* hdf5_abi/H5PLextern.h — plugin entry-point types (HDF5 H5PLextern.h / H5PLpublic.h). */
#ifndef DRICE_H5PLEXTERN_ABI_H
#define DRICE_H5PLEXTERN_ABI_H
#include "hdf5.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef enum H5PL_type_t {
    H5PL_TYPE_ERROR  = -1,
    H5PL_TYPE_FILTER = 0,
    H5PL_TYPE_VOL    = 1,
    H5PL_TYPE_VFD    = 2,
    H5PL_TYPE_NONE   = 3
} H5PL_type_t;
__attribute__((visibility("default"))) H5PL_type_t H5PLget_plugin_type(void);
__attribute__((visibility("default"))) const void *H5PLget_plugin_info(void);
#ifdef __cplusplus
}
#endif
#endif
