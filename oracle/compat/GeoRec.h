// oracle/compat/GeoRec.h — TEST INFRASTRUCTURE ONLY.  ZJU::GeoRec is a proprietary reconstruction
// library the reference links (Reconstruction/GeometryRec.h:4); only the type name is needed for
// Processor/Processor.cpp to parse.  Reconstruction is outside the compared path.
#ifndef MVS_COMPAT_GEOREC_H
#define MVS_COMPAT_GEOREC_H
namespace ZJU {
class GeoRec {};
}  // namespace ZJU
#endif
