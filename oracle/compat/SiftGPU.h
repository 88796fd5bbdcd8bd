// oracle/compat/SiftGPU.h — TEST INFRASTRUCTURE ONLY.  The one SiftGPU name the reference's headers
// mention (FeatureProc/FeatureProc.h), so that Processor/Processor.cpp parses; SiftGPU is not in this
// image and feature detection is outside the compared path.
#ifndef MVS_COMPAT_SIFTGPU_H
#define MVS_COMPAT_SIFTGPU_H
class SiftGPU {
 public:
  struct SiftKeypoint {
    float x, y, s, o;
  };
};
#endif
