// TEST INFRASTRUCTURE ONLY: empty stand-in (see cuda_runtime.h in this directory)
#pragma once
