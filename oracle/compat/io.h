// oracle/compat/io.h — TEST INFRASTRUCTURE ONLY: MSVC <io.h> names used by the reference's
// Common/Utils.h:92 (CreateDir), mapped to POSIX.
#ifndef MVS_COMPAT_IO_H
#define MVS_COMPAT_IO_H
#include <unistd.h>
static inline int _access(const char* path, int mode) { return access(path, mode); }
#endif
