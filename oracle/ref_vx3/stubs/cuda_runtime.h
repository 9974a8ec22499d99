// TEST INFRASTRUCTURE — see vxhost.h (force-included first); the reference's "#include <cuda_runtime.h>" lands here.
#pragma once
#include "vxhost.h"
