// Minimal XML reader for scene.xml (the reference uses pugixml, which is not vendored:
// src/core/Scene.cpp:3).  Elements, attributes, comments, <?...?> declarations and
// self-closing tags; no entities beyond the five predefined ones, no CDATA, no DTD.
#pragma once
#include <cstdlib>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace zillum {

class XmlNode {
public:
    XmlNode() {}
    explicit operator bool() const { return mData != nullptr; }
    const std::string& name() const { static std::string e; return mData ? mData->name : e; }
    // first child element with this name, or an empty node
    XmlNode child(const std::string& n) const {
        if (mData) for (auto& c : mData->children) if (c->name == n) return XmlNode(c);
        return XmlNode();
    }
    std::vector<XmlNode> children() const {
        std::vector<XmlNode> out;
        if (mData) for (auto& c : mData->children) out.push_back(XmlNode(c));
        return out;
    }
    bool hasAttribute(const std::string& n) const {
        if (mData) for (auto& a : mData->attrs) if (a.first == n) return true;
        return false;
    }
    // attribute value, "" when absent (pugi::xml_attribute::as_string semantics)
    std::string attribute(const std::string& n) const {
        if (mData) for (auto& a : mData->attrs) if (a.first == n) return a.second;
        return std::string();
    }
    int attributeInt(const std::string& n) const { return std::atoi(attribute(n).c_str()); }
    float attributeFloat(const std::string& n) const { return (float)std::atof(attribute(n).c_str()); }

    static XmlNode parseFile(const std::string& path, std::string* error = nullptr);
    static XmlNode parseString(const std::string& text, std::string* error = nullptr);

private:
    struct Data {
        std::string name;
        std::vector<std::pair<std::string, std::string>> attrs;
        std::vector<std::shared_ptr<Data>> children;
    };
    explicit XmlNode(std::shared_ptr<Data> d) : mData(std::move(d)) {}
    std::shared_ptr<Data> mData;
    friend class XmlParser;
};

}  // namespace zillum
