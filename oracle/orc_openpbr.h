// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.h header).
// OpenPBR closure -- placeholder until the layered restatement of open_pbr_surface.mtlx lands:
// diffuse base only.
#pragma once
#include "orc_bsdf.h"
namespace orc {
inline void openpbr_sample(const Material& m, BsdfSampleData& d, const State& s)
{
  Material dm; dm.closure = GTLC_CLOSURE_DIFFUSE;
  float w = m.p[GTLC_OPBR_BASE_WEIGHT];
  dm.p[0] = m.p[GTLC_OPBR_BASE_COLOR_R] * w; dm.p[1] = m.p[GTLC_OPBR_BASE_COLOR_G] * w; dm.p[2] = m.p[GTLC_OPBR_BASE_COLOR_B] * w;
  diffuse_sample(dm, d, s);
}
inline void openpbr_evaluate(const Material& m, BsdfEvaluateData& d, const State& s)
{
  Material dm; dm.closure = GTLC_CLOSURE_DIFFUSE;
  float w = m.p[GTLC_OPBR_BASE_WEIGHT];
  dm.p[0] = m.p[GTLC_OPBR_BASE_COLOR_R] * w; dm.p[1] = m.p[GTLC_OPBR_BASE_COLOR_G] * w; dm.p[2] = m.p[GTLC_OPBR_BASE_COLOR_B] * w;
  diffuse_evaluate(dm, d, s);
}
}
