#pragma once
#include "../../include/csts_b200.h"
