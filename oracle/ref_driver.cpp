// oracle/ref_driver.cpp -- C entry points around the reference's OWN hyperpose::parser::paf
// (compiled verbatim from /root/reference/src/paf.cpp over oracle/shim by oracle/Makefile).
// TEST INFRASTRUCTURE ONLY: used to validate paf_oracle.c and as the "reference" CPU baseline.
#include <cstring>
#include <hyperpose/operator/parser/paf.hpp>
#include "paf_oracle.h"

namespace hyperpose {
// src/data.cpp is not compiled into _ref (it needs cv::resize/copyMakeBorder on u8 images);
// the only symbol of it the parser needs is this member-wise constructor (data.hpp:22).
feature_map_t::feature_map_t(std::string name, std::unique_ptr<char[]>&& tensor, std::vector<int> shape)
    : m_name(std::move(name)), m_data(std::move(tensor)), m_shape(std::move(shape))
{
}
}

static hyperpose::feature_map_t make_map(const char* name, const float* src, int c, int h, int w)
{
    const size_t bytes = sizeof(float) * (size_t)c * h * w;
    std::unique_ptr<char[]> buf(new char[bytes]);
    std::memcpy(buf.get(), src, bytes);
    return hyperpose::feature_map_t(name, std::move(buf), { c, h, w });
}

extern "C" {
void* ref_paf_create(float conf_thresh, float paf_thresh, int res_w, int res_h)
{
    return new hyperpose::parser::paf(conf_thresh, paf_thresh, cv::Size(res_w, res_h));
}
void ref_paf_destroy(void* p) { delete static_cast<hyperpose::parser::paf*>(p); }
// one frame; returns number of humans written (<= cap), or -2 when cap is too small
int ref_paf_process(void* p, const float* conf, const float* paf, int c_conf, int c_paf, int H, int W,
    orc_human* out, int cap)
{
    auto* parser = static_cast<hyperpose::parser::paf*>(p);
    const auto cm = make_map("conf", conf, c_conf, H, W);
    const auto pm = make_map("paf", paf, c_paf, H, W);
    const std::vector<hyperpose::human_t> hs = parser->process(cm, pm);
    if ((int)hs.size() > cap) return -2;
    for (size_t i = 0; i < hs.size(); ++i) {
        std::memset(&out[i], 0, sizeof(orc_human));
        out[i].score = hs[i].score;
        for (int k = 0; k < ORC_N_PARTS; ++k) {
            out[i].parts[k].has_value = hs[i].parts[k].has_value ? 1 : 0;
            out[i].parts[k].x = hs[i].parts[k].x;
            out[i].parts[k].y = hs[i].parts[k].y;
            out[i].parts[k].score = hs[i].parts[k].score;
        }
    }
    return (int)hs.size();
}
}
