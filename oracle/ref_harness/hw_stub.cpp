// Test infrastructure (oracle/Makefile.ref). The BSDF plugins create their constant textures from the reference's hardware library
// (ConstantSpectrumTexture / ConstantFloatTexture, src/libhw/basicshader.cpp, compiled in place). That file refers to two entry points
// of the OpenGL preview renderer (src/libhw/renderer.cpp needs GL / GLX headers, absent here) from its GLSL shader classes, which only
// the interactive preview instantiates. These stand-ins satisfy the linker and fail loudly if anything ever calls them.
#include <mitsuba/hw/renderer.h>

MTS_NAMESPACE_BEGIN

Shader *Renderer::registerShaderForResource(const HWResource *) {
    SLog(EError, "oracle/_ref: the OpenGL preview renderer is not part of this build of the reference");
    return NULL;
}

void Renderer::unregisterShaderForResource(const HWResource *) {}

MTS_NAMESPACE_END
