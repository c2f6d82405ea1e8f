// oracle/ref_pifpaf_driver.cpp -- C entry point around the reference's OWN hyperpose::parser::pifpaf
// (src/pifpaf.cpp + src/pifpaf_decoder/*.cpp compiled verbatim from /root/reference by oracle/Makefile).
// TEST INFRASTRUCTURE ONLY: the checker / CPU baseline of the PifPaf decode path.
#include <cstring>
#include <hyperpose/operator/parser/pifpaf.hpp>
#include "paf_oracle.h"

namespace hyperpose {
feature_map_t::feature_map_t(std::string name, std::unique_ptr<char[]>&& tensor, std::vector<int> shape)
    : m_name(std::move(name)), m_data(std::move(tensor)), m_shape(std::move(shape))
{
}
}

static hyperpose::feature_map_t make_map4(const char* name, const float* src, int d0, int d1, int h, int w)
{
    const size_t bytes = sizeof(float) * (size_t)d0 * d1 * h * w;
    std::unique_ptr<char[]> buf(new char[bytes]);
    std::memcpy(buf.get(), src, bytes);
    return hyperpose::feature_map_t(name, std::move(buf), { d0, d1, h, w });
}

extern "C" {
// pif [17,5,h,w], paf [19,9,h,w] (feature-cell units).  NOTE the reference's argument order quirk:
// pifpaf::process is DEFINED as (paf, pif) (src/pifpaf.cpp:7) although the header declares (pif, paf); callers pass
// packet[0], packet[1] = outputs sorted by name.  This driver passes the tensors in the DEFINITION's order.
int ref_pifpaf_process(const float* pif, const float* paf, int h, int w, int net_h, int net_w, float keypoint_thresh,
    orc_human* out, int cap)
{
    hyperpose::parser::pifpaf parser(net_h, net_w, keypoint_thresh);
    const auto pafm = make_map4("paf", paf, 19, 9, h, w);
    const auto pifm = make_map4("pif", pif, 17, 5, h, w);
    const std::vector<hyperpose::human_t> hs = parser.process(pafm, pifm);
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
