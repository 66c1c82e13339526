// spectral_tc.cuh - tcgen05 / TMEM tensor-core kernels (sm_100a only).  Filled in once the SIMT
// path is parity-green; see DESIGN.md.
#pragma once
#include "spectral_common.cuh"
