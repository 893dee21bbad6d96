// Stand-in for <cuda_runtime.h> when libgfk's kernel sources are compiled for the SIMT host emulator.
#pragma once
#include "../simt.h"
