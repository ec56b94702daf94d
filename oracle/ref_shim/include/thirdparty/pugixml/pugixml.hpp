// ORACLE — test infrastructure only.  Part of the recipe that builds oracle/_ref.
// pugixml.hpp — stand-in for pugixml (absent from the image and from the reference tree): the read-only DOM calls
// that src/core/Scene.cpp:7-127 and src/core/MaterialLoader.cpp:5-69 make (load_file / load_string, child, first_child,
// next_sibling, attribute, as_string / as_int / as_float, truth test), over a small own parser.  Missing nodes and
// attributes behave like pugixml's null handles: every call on them is valid and yields "" / 0.
#pragma once
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

namespace pugi {

struct xml_node_data {
    std::string name;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<std::unique_ptr<xml_node_data>> children;
    xml_node_data* parent = nullptr;
    size_t indexInParent = 0;
};

class xml_attribute {
public:
    xml_attribute(const std::string* v = nullptr) : mValue(v) {}
    const char* as_string(const char* def = "") const { return mValue ? mValue->c_str() : def; }
    int as_int(int def = 0) const { return mValue ? (int)std::strtol(mValue->c_str(), nullptr, 10) : def; }
    float as_float(float def = 0.0f) const { return mValue ? (float)std::strtod(mValue->c_str(), nullptr) : def; }
    explicit operator bool() const { return mValue != nullptr; }
private:
    const std::string* mValue;
};

class xml_node {
public:
    xml_node(const xml_node_data* d = nullptr) : mData(d) {}
    explicit operator bool() const { return mData != nullptr; }
    bool operator!() const { return mData == nullptr; }
    const char* name() const { return mData ? mData->name.c_str() : ""; }
    xml_node child(const char* n) const {
        if (mData) for (auto& c : mData->children) if (c->name == n) return xml_node(c.get());
        return xml_node();
    }
    xml_node first_child() const { return (mData && !mData->children.empty()) ? xml_node(mData->children[0].get()) : xml_node(); }
    xml_node next_sibling() const {
        if (!mData || !mData->parent) return xml_node();
        size_t i = mData->indexInParent + 1;
        return i < mData->parent->children.size() ? xml_node(mData->parent->children[i].get()) : xml_node();
    }
    xml_attribute attribute(const char* n) const {
        if (mData) for (auto& a : mData->attrs) if (a.first == n) return xml_attribute(&a.second);
        return xml_attribute();
    }
protected:
    const xml_node_data* mData;
};

struct xml_parse_result { bool ok; explicit operator bool() const { return ok; } };

class xml_document : public xml_node {
public:
    xml_document() : mRoot(new xml_node_data) { mData = mRoot.get(); }
    xml_parse_result load_file(const char* path) {
        std::ifstream f(path);
        if (!f.is_open()) return {false};
        std::stringstream ss; ss << f.rdbuf();
        return load_string(ss.str().c_str());
    }
    xml_parse_result load_string(const char* text) {
        mRoot.reset(new xml_node_data); mData = mRoot.get();
        const std::string s(text);
        size_t i = 0;
        xml_node_data* cur = mRoot.get();
        while (i < s.size()) {
            size_t lt = s.find('<', i);
            if (lt == std::string::npos) break;
            if (s.compare(lt, 4, "<!--") == 0) { size_t e = s.find("-->", lt); if (e == std::string::npos) break; i = e + 3; continue; }
            if (s.compare(lt, 2, "<?") == 0) { size_t e = s.find("?>", lt); if (e == std::string::npos) break; i = e + 2; continue; }
            if (s.compare(lt, 2, "</") == 0) { size_t e = s.find('>', lt); if (e == std::string::npos) break; if (cur->parent) cur = cur->parent; i = e + 1; continue; }
            size_t p = lt + 1;
            auto isName = [](char c) { return std::isalnum((unsigned char)c) || c == '_' || c == '-' || c == ':' || c == '.'; };
            size_t q = p; while (q < s.size() && isName(s[q])) q++;
            std::unique_ptr<xml_node_data> node(new xml_node_data);
            node->name = s.substr(p, q - p); node->parent = cur; node->indexInParent = cur->children.size();
            bool selfClose = false;
            for (;;) {
                while (q < s.size() && std::isspace((unsigned char)s[q])) q++;
                if (q >= s.size()) break;
                if (s[q] == '/') { selfClose = true; q = s.find('>', q); break; }
                if (s[q] == '>') break;
                size_t a = q; while (q < s.size() && isName(s[q])) q++;
                std::string an = s.substr(a, q - a);
                while (q < s.size() && (std::isspace((unsigned char)s[q]) || s[q] == '=')) q++;
                if (q >= s.size() || (s[q] != '"' && s[q] != '\'')) { q++; continue; }
                char quote = s[q++]; size_t v = q; while (q < s.size() && s[q] != quote) q++;
                node->attrs.emplace_back(an, unescape(s.substr(v, q - v)));
                q++;
            }
            if (q == std::string::npos || q >= s.size()) break;
            xml_node_data* raw = node.get();
            cur->children.push_back(std::move(node));
            if (!selfClose) cur = raw;
            i = q + 1;
        }
        return {true};
    }
private:
    static std::string unescape(const std::string& in) {      // the five predefined entities, as pugixml's parse_escapes does
        static const char* const ent[5][2] = {{"&amp;", "&"}, {"&lt;", "<"}, {"&gt;", ">"}, {"&quot;", "\""}, {"&apos;", "'"}};
        std::string out;
        for (size_t i = 0; i < in.size();) {
            bool hit = false;
            if (in[i] == '&')
                for (auto& e : ent) { size_t n = std::strlen(e[0]); if (in.compare(i, n, e[0]) == 0) { out += e[1]; i += n; hit = true; break; } }
            if (!hit) out += in[i++];
        }
        return out;
    }
    std::unique_ptr<xml_node_data> mRoot;
};

}  // namespace pugi
