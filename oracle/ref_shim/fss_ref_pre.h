/* Force-included (-include) in front of every translation unit of oracle/_ref.
 * Pulls the standard headers in FIRST, then (unless FSS_REF_LIBC_RAND) routes every later
 * `rand()` to the keyed RNG of include/fss_rng.h with site = FSS_REF_SITE_BASE + __LINE__.
 * Test infrastructure; not part of the product. */
#ifndef FSS_REF_PRE_H
#define FSS_REF_PRE_H
#include <math.h> /* like box2d/b2_math.h in the game: brings the float overloads of abs() into the global namespace */
#include <stdlib.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <future>
#include <iostream>
#include <string>
#include <unordered_map>
#include <vector>

int fss_ref_rand(int site);
void fss_ref_ctx(int tick_ct, int iter, int x, int y);

#ifndef FSS_REF_SITE_BASE
#define FSS_REF_SITE_BASE 0
#endif
#ifndef FSS_REF_LIBC_RAND
#define rand() fss_ref_rand(FSS_REF_SITE_BASE + __LINE__)
#endif
#endif
