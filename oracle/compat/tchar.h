// oracle/compat/tchar.h — TEST INFRASTRUCTURE ONLY: the ANSI form of the _T() macro
// (Common/Utils.h:60).
#ifndef MVS_COMPAT_TCHAR_H
#define MVS_COMPAT_TCHAR_H
#define _T(x) x
#endif
