// Flash-attention entry points (filled in by csrc/attn/*.cu).
#pragma once
#include <cuda_runtime.h>
namespace tb {}
