// oracle/compat/msvc_prelude.h — TEST INFRASTRUCTURE ONLY.  Force-included (-include) when the
// reference's first-party sources are compiled unmodified with g++ (oracle/Makefile target `ref`):
// the MSVC CRT names they use without including a Windows header (SetUtils/PointSetUtils.cpp:53-58
// `__min/__max`).
#ifndef MVS_COMPAT_MSVC_PRELUDE_H
#define MVS_COMPAT_MSVC_PRELUDE_H
#include <math.h>
#include <stdlib.h>
#ifndef __min
#define __min(a, b) (((a) < (b)) ? (a) : (b))
#endif
#ifndef __max
#define __max(a, b) (((a) > (b)) ? (a) : (b))
#endif
#endif
