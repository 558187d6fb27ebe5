// Internal (non-exported) interfaces between the translation units of libshb200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
