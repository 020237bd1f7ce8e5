// loco_proto.h -- reader/writer for the reference's on-disk format.
//
// Replaces PrecomputedParametricShapeConverter::loadFromProto / outputToProto
// (LCAdaptiveSampling/LCProtoConverter.h:11-21, .cpp:14-33) and the ProtoConverter<T> file I/O
// (LCFunction/LCShapeInfoProtoConverter.h:11-80).  The reference links google/protobuf v3.1.0
// (README.md:21-32), which is not available here, so the proto2 wire format of
// LCFunction/PrecomputedParametricShape.proto is decoded / encoded by hand.
#pragma once
#include "loco_graph.h"

namespace loco {

// binary wire format (ascii=false) or protobuf text format (ascii=true)
Status read_proto_file(const std::string &path, bool ascii, Graph *out);
Status write_proto_file(const Graph &g, const std::string &path, bool ascii);

Status decode_proto(const uint8_t *data, size_t size, Graph *out);
Status encode_proto(const Graph &g, std::string *out);
Status decode_proto_text(const std::string &text, Graph *out);
Status encode_proto_text(const Graph &g, std::string *out);

} // namespace loco
