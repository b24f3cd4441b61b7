// opencv2/highgui/highgui.hpp — COMPAT LAYER: the reference includes it but uses nothing from it.
#pragma once
#include <opencv2/core/core.hpp>
