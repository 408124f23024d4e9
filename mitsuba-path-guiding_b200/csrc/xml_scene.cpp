// xml_scene.cpp -- Mitsuba 0.6 XML subset reader (scenehandler.cpp semantics).
#include "host_scene.h"

namespace pg {
bool loadSceneXml(const char *, const char *const *, HostScene &, std::string &err) {
    err = "XML loader not built yet";
    return false;
}
}  // namespace pg
