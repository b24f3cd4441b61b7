// opencv2/core.hpp — COMPAT LAYER (see core/core.hpp).
#pragma once
#include <opencv2/core/core.hpp>
