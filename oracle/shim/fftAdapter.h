// Build-side shim (test infrastructure, not product code).
// The reference's include/grid.h:7 says #include "fftAdapter.h" but the file on disk is
// include/FFTAdapter.h (case-insensitive Windows file system).  This header bridges the case.
#pragma once
#include "FFTAdapter.h"
