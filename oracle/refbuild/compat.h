/* Force-included when compiling the reference with g++ 13: supplies headers the
 * reference relies on transitively with older compilers, and disambiguates
 * abs(size_t) (Algorithm/KmerOverlaps.cpp:209). Test infrastructure only. */
#ifdef __cplusplus
#include <cstddef>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <string>
#include <limits>
#include <array>
#include <memory>
#include <functional>
inline size_t abs(size_t x) { return x; }
#endif
