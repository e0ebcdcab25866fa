// TEST INFRASTRUCTURE ONLY: Win32 names for the nvcc build of the reference (oracle/build_ref.py,
// build_cuda).  Kept in its own directory so that the host-emulation cuda_runtime.h of oracle/shim/
// is NOT on the include path of that build -- it uses the real CUDA runtime.
#pragma once
#include "../shim/windows.h"
