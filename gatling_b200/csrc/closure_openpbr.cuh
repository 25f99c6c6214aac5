// OpenPBR closure -- placeholder until the layered restatement of
// /root/reference/src/gi/mtlx/open_pbr_surface.mtlx:99-635 lands: diffuse base only.
#pragma once
namespace gtlc {
GTLC_D void openpbr_sample(const float* p, BsdfSample& d, const ShadingState& s)
{
  float w = p[GTLC_OPBR_BASE_WEIGHT];
  float dp[4] = { p[GTLC_OPBR_BASE_COLOR_R] * w, p[GTLC_OPBR_BASE_COLOR_G] * w, p[GTLC_OPBR_BASE_COLOR_B] * w, 0.0f };
  diffuse_sample(dp, d, s);
}
GTLC_D void openpbr_evaluate(const float* p, BsdfEval& d, const ShadingState& s)
{
  float w = p[GTLC_OPBR_BASE_WEIGHT];
  float dp[4] = { p[GTLC_OPBR_BASE_COLOR_R] * w, p[GTLC_OPBR_BASE_COLOR_G] * w, p[GTLC_OPBR_BASE_COLOR_B] * w, 0.0f };
  diffuse_evaluate(dp, d, s);
}
}
