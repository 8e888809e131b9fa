// oracle/ref_transfer_shim.h -- TEST INFRASTRUCTURE.  Force-included when the reference's transfer.cpp is compiled for the
// oracle (oracle/Makefile, target ref_transfer).  transfer.cpp splits its two colour-space conversions over Win32 threads
// (CreateThread / WaitForMultipleObjects, transfer.cpp:42-121); every thread converts its own band of rows in place and
// the bands are disjoint, so running each "thread" to completion inside CreateThread gives the same bytes.  Nothing
// arithmetic is touched.
#pragma once
#include <cstddef>
typedef void* HANDLE;
typedef unsigned long DWORD;
typedef void* LPVOID;
typedef DWORD (*LPTHREAD_START_ROUTINE)(LPVOID);
#ifndef TRUE
#define TRUE 1
#endif
#define INFINITE 0xFFFFFFFFu
static inline HANDLE CreateThread(void*, size_t, LPTHREAD_START_ROUTINE fn, LPVOID arg, DWORD, DWORD*) {
    fn(arg);                       // the band is converted before CreateThread returns
    return (HANDLE)0;
}
static inline DWORD WaitForMultipleObjects(DWORD, const HANDLE*, int, DWORD) { return 0; }
static inline int CloseHandle(HANDLE) { return 1; }
