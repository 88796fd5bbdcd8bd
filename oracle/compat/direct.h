// oracle/compat/direct.h — TEST INFRASTRUCTURE ONLY: MSVC <direct.h> names used by the reference's
// Common/Utils.h:94 (CreateDir), mapped to POSIX.
#ifndef MVS_COMPAT_DIRECT_H
#define MVS_COMPAT_DIRECT_H
#include <sys/stat.h>
#include <sys/types.h>
static inline int _mkdir(const char* path) { return mkdir(path, 0777); }
#endif
