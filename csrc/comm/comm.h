// Symmetric-memory communication runtime entry points (filled in by csrc/comm/*.cu).
#pragma once
#include <cuda_runtime.h>
namespace tb {}
