// Test infrastructure (oracle/Makefile.ref). Stand-in for the reference's src/libcore/fmtconv.cpp, which instantiates its pixel-format
// converters with Boost.MPL and cannot be compiled in this image. Bitmap links against these three entry points of the
// FormatConverter interface (include/mitsuba/core/bitmap.h:1387-1462); the registry stays empty, so any attempt to convert a
// bitmap between component formats (developing a film to a file, tonemapping, ...) fails loudly. Nothing on the path the
// harness drives -- ray casts, Li, BSDF / emitter / medium queries, ImageBlock::put -- converts a bitmap.
#include <mitsuba/core/bitmap.h>

MTS_NAMESPACE_BEGIN

FormatConverter::ConverterMap FormatConverter::m_converters;

void FormatConverter::staticInitialization() {}

void FormatConverter::staticShutdown() {}

const FormatConverter *FormatConverter::getInstance(Conversion) {
    SLog(EError, "oracle/_ref: pixel-format conversion is not part of this build of the reference (no Boost.MPL in the image)");
    return NULL;
}

MTS_NAMESPACE_END
