// ORACLE -- TEST INFRASTRUCTURE ONLY.  BVH8 walker (filled in once the node layout is fixed).
#pragma once
