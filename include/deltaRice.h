/* deltaRice.h — drop-in replacement of the reference's public header
 * (reference src/deltaRice.h:1-17): same macro, same three declarations, same filter id.
 * Programs that include the reference's deltaRice.h and call
 * deltarice_register_h5filter() (reference examples/testCode.c:38) compile unchanged
 * against this header and link against libh5deltarice_b200.so instead.
 *
 * Build with -DDRICE_USE_SYSTEM_HDF5 where real HDF5 headers exist; otherwise the
 * vendored ABI slice in include/hdf5_abi/ is used. */
#ifndef DELTARICE_H5FILTER_H
#define DELTARICE_H5FILTER_H

#define H5Z_class_t_vers 2
#ifdef DRICE_USE_SYSTEM_HDF5
#include "hdf5.h"
#else
#include "hdf5_abi/hdf5.h"
#endif

#define H5Z_FILTER_DELTARICE 32025            /* reference src/deltaRice.h:7 */
typedef unsigned long long int superint;      /* reference src/deltaRice.h:8 */

#if defined(__GNUC__)
#define DRICE_H5_API __attribute__((visibility("default")))
#else
#define DRICE_H5_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* reference src/deltaRice.h:10 / src/deltaRice.c:19-28 */
extern DRICE_H5_API H5Z_class_t H5Z_DELTARICE[1];

/* reference src/deltaRice.h:13 / src/deltaRice.c:468-490.
 * Same contract towards libhdf5: *buf is malloc-family memory holding nbytes valid bytes;
 * on success the filter frees it, stores a new malloc'ed buffer in *buf, its size in
 * *buf_size and returns that size.  On failure it returns 0 and leaves *buf untouched
 * (HDF5's convention; the reference returns (size_t)-1). */
DRICE_H5_API size_t H5Z_filter_deltarice(unsigned flags, size_t cd_nelmts, const unsigned cd_values[],
                            size_t nbytes, size_t *buf_size, void **buf);

/* reference src/deltaRice.h:15 / src/deltaRice.c:494-501: H5Zregister(H5Z_DELTARICE).
 * libhdf5 is resolved from the running process at call time (no link-time dependency),
 * returns < 0 when no libhdf5 is loaded or registration fails. */
DRICE_H5_API int deltarice_register_h5filter(void);

#ifdef __cplusplus
}
#endif
#endif /* DELTARICE_H5FILTER_H */
