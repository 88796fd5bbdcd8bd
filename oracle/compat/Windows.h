// oracle/compat/Windows.h — TEST INFRASTRUCTURE ONLY.  The handful of Win32 names the reference's
// Common/Utils.h:53-80 (ScanNSortDirectory) mentions, so the header parses with g++.  Directory
// scanning is not on the hot path: FindFirstFile always reports "nothing found".
#ifndef MVS_COMPAT_WINDOWS_H
#define MVS_COMPAT_WINDOWS_H
#include "msvc_prelude.h"
typedef void* HANDLE;
typedef unsigned long DWORD;
#define INVALID_HANDLE_VALUE ((HANDLE)(long)-1)
#define FILE_ATTRIBUTE_DIRECTORY 0x10
struct WIN32_FIND_DATA {
  DWORD dwFileAttributes;
  char cFileName[260];
};
static inline HANDLE FindFirstFile(const char*, WIN32_FIND_DATA*) { return INVALID_HANDLE_VALUE; }
static inline int FindNextFile(HANDLE, WIN32_FIND_DATA*) { return 0; }
static inline int FindClose(HANDLE) { return 1; }
// Processor/Processor.cpp:925-1014 renames result files
#include <stdio.h>
#define MOVEFILE_REPLACE_EXISTING 0x1
static inline int MoveFileEx(const char* from, const char* to, DWORD) { return rename(from, to) == 0; }
#endif
