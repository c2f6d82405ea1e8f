// shim: <opencv2/core/utils/filesystem.hpp> -- cv::utils::fs::glob as examples/utils.cpp:23-52 uses it:
// glob(directory, "*.ext", out): the files of `directory` (not recursive) whose names end in .ext, sorted.
#pragma once
#include <algorithm>
#include <filesystem>
#include <string>
#include <vector>

#include "../../opencv.hpp"

namespace cv { namespace utils { namespace fs {
inline void glob(const cv::String& directory, const cv::String& pattern, std::vector<cv::String>& result, bool /*recursive*/ = false, bool /*includeDirectories*/ = false)
{
    namespace sfs = std::filesystem;
    const std::string suffix = (!pattern.empty() && pattern[0] == '*') ? pattern.substr(1) : pattern;
    std::error_code ec;
    if (!sfs::is_directory(directory, ec)) return;
    for (const auto& e : sfs::directory_iterator(directory, ec)) {
        if (!e.is_regular_file()) continue;
        const std::string name = e.path().filename().string();
        if (name.size() >= suffix.size() && name.compare(name.size() - suffix.size(), suffix.size(), suffix) == 0) result.push_back(e.path().string());
    }
    std::sort(result.begin(), result.end());
}
}}} // namespace cv::utils::fs
