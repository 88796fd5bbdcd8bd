// oracle/compat/msvc_prelude.h — TEST INFRASTRUCTURE ONLY.  Force-included (-include) when the
// reference's first-party sources are compiled unmodified with g++ (oracle/Makefile target `ref`):
// the MSVC CRT names they use without including a Windows header (SetUtils/PointSetUtils.cpp:53-58
// `__min/__max`).
#ifndef MVS_COMPAT_MSVC_PRELUDE_H
#define MVS_COMPAT_MSVC_PRELUDE_H
#include <math.h>
#include <stdlib.h>
#ifdef __cplusplus
// headers MSVC's own standard headers pull in transitively (Alignment.cpp uses std::numeric_limits,
// PlyObj.cpp std::find, without including <limits> / <algorithm>)
#include <algorithm>
#include <iomanip>
#include <limits>
#include <cstdio>
#include <cstring>
#include <ctime>
// MSVC binds a temporary to a non-const reference parameter's default argument
// (Alignment/Alignment.h:31-36, PlyObj/PlyObj.h:89-90,237-238,293-294,345-346).  g++ has no switch
// for that, so `make ref` passes those two HEADERS through one sed rule that rewrites
// `T &x = T()` into `T &x = msvc_default_arg<T >()`: a fresh, empty object per call, as on MSVC.
template <class T>
inline T& msvc_default_arg() {
  static thread_local T t;
  t = T();
  return t;
}
#endif
// the "secure" CRT variants used with plain numeric formats only (PlyObj/PlyObj.cpp:46-67)
#ifndef sscanf_s
#define sscanf_s sscanf
#endif
#ifdef __cplusplus
// sprintf_s(char (&)[N], fmt, ...) — the array-size-deducing overload Processor/Processor.cpp uses
template <size_t N, class... A>
inline int sprintf_s(char (&buf)[N], const char* fmt, A... a) {
  return snprintf(buf, N, fmt, a...);
}
template <class... A>
inline int sprintf_s(char* buf, size_t n, const char* fmt, A... a) {
  return snprintf(buf, n, fmt, a...);
}
#endif
#ifndef __min
#define __min(a, b) (((a) < (b)) ? (a) : (b))
#endif
#ifndef __max
#define __max(a, b) (((a) > (b)) ? (a) : (b))
#endif
#endif
