// shim: <opencv2/core/mat.hpp> (examples/utils.hpp:3) -> the one shim header
#pragma once
#include "../opencv.hpp"
