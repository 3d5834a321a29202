/* Forced into the reference build with -include.  The reference's OpenMP encoder
 * allocates 2*nbytes+numWaves+1 BYTES (src/deltaRice.c:412) but stages wave i at WORD
 * i*L+i+1 (:421), so it overruns its own heap block for many-wave or incompressible
 * chunks.  Over-allocating keeps the unmodified source usable as
 * an oracle without touching its arithmetic. */
#include <stdlib.h>
static inline void *drice_slack_malloc(size_t n) { return malloc(4 * n + 65536); }
#define malloc(x) drice_slack_malloc(x)
