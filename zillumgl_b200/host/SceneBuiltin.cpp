// Built-in benchmark scenes, expressed in the reference's own scene.xml dialect
// (res/scene.xml:1-49) with "builtin:" model paths, so they go through the same loader as
// a user scene.  C1 "default" is the commented-out block of res/scene.xml:19-44 made
// concrete; the others are the synthetic stand-ins of SURVEY.md §8(d).
#include <cstdio>
#include <sstream>
#include "Scene.h"

namespace zillum {

std::string Scene::builtinXml(const std::string& nameIn, int w, int h) {
    std::string name = nameIn, query;
    size_t q = name.find('?');
    if (q != std::string::npos) { query = name.substr(q); name = name.substr(0, q); }
    std::ostringstream x;
    x << "<?xml version=\"1.0\"?>\n<scene name=\"" << name << "\">\n";
    auto head = [&](const char* integ, const char* sampler, const char* pos, const char* angle, float fov, float lens, float focal) {
        x << "  <integrator type=\"" << integ << "\"><maxBounce value=\"4\"/><size width=\"" << w << "\" height=\"" << h << "\"/></integrator>\n"
          << "  <sampler type=\"" << sampler << "\"/>\n"
          << "  <camera type=\"thinLens\"><position value=\"" << pos << "\"/><angle value=\"" << angle << "\"/><fov value=\"" << fov
          << "\"/><lensRadius value=\"" << lens << "\"/><focalDistance value=\"" << focal << "\"/></camera>\n  <modelInstances>\n";
    };
    if (name == "default") {
        head("path", "sobol", "0 -8 3", "0 0 0", 45, 0, 1);
        x << "    <modelInstance path=\"builtin:square\" name=\"square\" type=\"object\"><transform translate=\"0 0 0\" scale=\"100 100 1\" rotate=\"0 0 0\"/><material type=\"default\"/></modelInstance>\n"
             "    <modelInstance path=\"builtin:teapotBody\" name=\"teapotBody\" type=\"object\"><transform translate=\"0 0 0\" scale=\"1 1 1\" rotate=\"0 0 0\"/>"
             "<material type=\"metalWorkflow\"><albedo value=\"1 1 1\"/><metallic value=\"1\"/><roughness value=\"0.1\"/></material></modelInstance>\n"
             "    <modelInstance path=\"builtin:teapotCap\" name=\"teapotCap\" type=\"object\"><transform translate=\"0 0 0\" scale=\"1 1 1\" rotate=\"0 0 0\"/>"
             "<material type=\"dielectric\"><tint value=\"1 1 1\"/><ior value=\"1.5\"/><roughness value=\"0.0\"/></material></modelInstance>\n"
             "    <modelInstance path=\"builtin:square\" name=\"areaLight\" type=\"light\"><transform translate=\"0 0 10\" scale=\"2 2 1\" rotate=\"180 0 0\"/><radiance value=\"20 20 20\"/></modelInstance>\n"
             "  </modelInstances>\n";
    } else if (name == "cornell") {
        head("lightPath", "sobol", "0 -4.4 1", "0 0 0", 34, 0, 1);
        x << "    <modelInstance path=\"builtin:cornell\" name=\"room\" type=\"object\"><transform translate=\"0 0 0\" scale=\"1 1 1\" rotate=\"0 0 0\"/><material type=\"default\"/></modelInstance>\n"
             "    <modelInstance path=\"builtin:square\" name=\"ceilingLight\" type=\"light\"><transform translate=\"0 0 1.995\" scale=\"0.6 0.6 1\" rotate=\"180 0 0\"/><radiance value=\"12 10.8 8.4\"/></modelInstance>\n"
             "  </modelInstances>\n";
    } else if (name == "sponza" || name == "sponza_light") {
        head(name == "sponza" ? "path" : "triplePath", "sobol", "-17.5 0.6 2.2", "90 4 0", 55, 0, 1);
        x << "    <modelInstance path=\"builtin:sponza\" name=\"atrium\" type=\"object\"><transform translate=\"0 0 0\" scale=\"1 1 1\" rotate=\"0 0 0\"/><material type=\"default\"/></modelInstance>\n";
        if (name == "sponza_light")
            for (int i = 0; i < 4; i++)
                x << "    <modelInstance path=\"builtin:square\" name=\"lamp" << i << "\" type=\"light\"><transform translate=\"" << (-14 + 9 * i)
                  << " 0 9.5\" scale=\"1.5 1.5 1\" rotate=\"180 0 0\"/><radiance value=\"160 150 130\"/></modelInstance>\n";
        x << "  </modelInstances>\n  <envMap path=\"builtin:sky\"/>\n";
    } else if (name == "rungholt") {
        head("path", "sobol", "-430 -235 75", "62 -15 0", 50, 0, 1);
        x << "    <modelInstance path=\"builtin:rungholt" << query << "\" name=\"city\" type=\"object\"><transform translate=\"0 0 0\" scale=\"1 1 1\" rotate=\"0 0 0\"/><material type=\"default\"/></modelInstance>\n"
             "  </modelInstances>\n  <envMap path=\"builtin:sky\"/>\n";
    } else if (name == "rungholt_small") {     // test-sized city with the same generator
        head("path", "sobol", "-30 -40 25", "38 -28 0", 50, 0, 1);
        x << "    <modelInstance path=\"builtin:rungholt" << (query.empty() ? "?nx=64&amp;ny=48" : query) << "\" name=\"city\" type=\"object\"><transform translate=\"0 0 0\" scale=\"1 1 1\" rotate=\"0 0 0\"/><material type=\"default\"/></modelInstance>\n"
             "    <modelInstance path=\"builtin:square\" name=\"lamp\" type=\"light\"><transform translate=\"0 0 40\" scale=\"20 20 1\" rotate=\"180 0 0\"/><radiance value=\"9000 9000 8000\"/></modelInstance>\n"
             "  </modelInstances>\n  <envMap path=\"builtin:sky\"/>\n";
    } else {
        return std::string();
    }
    x << "</scene>\n";
    return x.str();
}

bool Scene::loadBuiltin(const std::string& name, int width, int height) {
    std::string xml = builtinXml(name, width, height);
    if (xml.empty()) { std::fprintf(stderr, "[Scene] unknown builtin scene '%s'\n", name.c_str()); return false; }
    return loadXmlText(xml);
}

}  // namespace zillum
