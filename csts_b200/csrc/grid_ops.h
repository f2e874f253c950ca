// The token-grid kernels' argument blocks (csts_pool_args, csts_wgrad_args) are part of the public
// C ABI: include/csts_b200.h.
#pragma once
#include "../../include/csts_b200.h"
