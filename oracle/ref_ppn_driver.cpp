// oracle/ref_ppn_driver.cpp -- C entry point around the reference's OWN hyperpose::parser::pose_proposal
// (src/pose_proposal.cpp compiled verbatim from /root/reference by oracle/Makefile).
// TEST INFRASTRUCTURE ONLY: the checker of the Pose Proposal Network parse path.
#include <cstring>
#include <hyperpose/operator/parser/proposal_network.hpp>
#include "paf_oracle.h"

namespace hyperpose {
feature_map_t::feature_map_t(std::string name, std::unique_ptr<char[]>&& tensor, std::vector<int> shape)
    : m_name(std::move(name)), m_data(std::move(tensor)), m_shape(std::move(shape))
{
}
}

static hyperpose::feature_map_t make_map(const char* name, const float* src, std::vector<int> shape)
{
    size_t n = 1;
    for (int d : shape) n *= (size_t)d;
    std::unique_ptr<char[]> buf(new char[n * sizeof(float)]);
    std::memcpy(buf.get(), src, n * sizeof(float));
    return hyperpose::feature_map_t(name, std::move(buf), std::move(shape));
}

extern "C" {
// conf_point/conf_iou/x/y/w/h: [K,gh,gw]; edge: [E,nh,nw,gh,gw] (src/pose_proposal.cpp:14-20,186)
int ref_ppn_process(const float* conf_point, const float* conf_iou, const float* x, const float* y, const float* w, const float* h,
    const float* edge, int K, int gh, int gw, int E, int nh, int nw, int net_w, int net_h,
    float point_thresh, float limb_thresh, float nms_thresh, orc_human* out, int cap)
{
    hyperpose::parser::pose_proposal parser(cv::Size(net_w, net_h), point_thresh, limb_thresh, nms_thresh);
    const auto hs = parser.process(make_map("conf_point", conf_point, { K, gh, gw }), make_map("conf_iou", conf_iou, { K, gh, gw }),
        make_map("x", x, { K, gh, gw }), make_map("y", y, { K, gh, gw }), make_map("w", w, { K, gh, gw }), make_map("h", h, { K, gh, gw }),
        make_map("edge", edge, { E, nh, nw, gh, gw }));
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
