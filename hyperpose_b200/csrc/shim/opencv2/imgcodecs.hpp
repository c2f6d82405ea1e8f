// shim: <opencv2/imgcodecs.hpp> (examples/utils.cpp:4) -> the one shim header (cv::imread / cv::imwrite on binary PPM)
#pragma once
#include "opencv.hpp"
