/* TEST INFRASTRUCTURE ONLY -- force-included (-include) in front of the UNMODIFIED
 * reference sources when oracle/build_ref.sh compiles them into oracle/_ref/.
 *
 * Purpose: the reference reads memory it never wrote (SURVEY.md finding 5):
 *   - Descriptor::I_desc borders   (src/common_includes/elas/descriptor.cpp:31, read by
 *     src/serial_includes/elas/elas.cpp:714-763)
 *   - adaptiveMean's D_tmp          (src/serial_includes/elas/elas.cpp:1308, read at :1448-1454)
 *   - Sobel temporaries' rows 0/H-1 (src/common_includes/elas/filter.cpp:417-418)
 * In the reference process those blocks come fresh from mmap and are therefore zero on the
 * first frame.  To make that the *defined* behaviour on every call, every malloc/_mm_malloc
 * of the reference translation units is turned into a zero-filling allocation.  Nothing else
 * about the reference changes.
 */
#ifndef ORACLE_REF_PRELUDE_H
#define ORACLE_REF_PRELUDE_H

/* pull in every header that declares malloc BEFORE the macros below exist */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <stdint.h>
#include <math.h>
#include <emmintrin.h>
#include <pmmintrin.h>
#include <mm_malloc.h>
#ifdef __cplusplus
#include <cstdlib>
#include <cmath>
#include <new>
#include <memory>
#include <string>
#include <vector>
#include <map>
#include <algorithm>
#include <iostream>
#include <fstream>
#include <sstream>
#endif

static inline void *oracle_ref_zmalloc(size_t n) { return calloc(1, n ? n : 1); }
static inline void *oracle_ref_mm_zmalloc(size_t n, size_t align) {
    void *p = _mm_malloc(n ? n : 1, align);
    if (p) memset(p, 0, n);
    return p;
}
#define malloc(n) oracle_ref_zmalloc(n)
#define _mm_malloc(n, a) oracle_ref_mm_zmalloc((n), (a))

#endif
