#include "Xml.h"
#include <cctype>
#include <fstream>
#include <sstream>

namespace zillum {

class XmlParser {
public:
    XmlParser(const std::string& s) : t(s) {}
    std::shared_ptr<XmlNode::Data> parseDocument(std::string* err) {
        auto root = std::make_shared<XmlNode::Data>();   // synthetic document node
        try {
            while (true) {
                skipMisc();
                if (p >= t.size()) break;
                if (t[p] != '<') fail("text outside of the root element");
                root->children.push_back(parseElement());
            }
        } catch (const std::string& e) {
            if (err) *err = e;
            return nullptr;
        }
        return root;
    }

private:
    const std::string& t;
    size_t p = 0;
    [[noreturn]] void fail(const std::string& m) { throw std::string("xml: " + m + " at offset " + std::to_string(p)); }
    bool starts(const char* s) const { return t.compare(p, std::char_traits<char>::length(s), s) == 0; }
    void skipWs() { while (p < t.size() && std::isspace((unsigned char)t[p])) p++; }
    void skipMisc() {   // whitespace, comments, declarations, doctype
        while (true) {
            skipWs();
            if (starts("<!--")) { size_t e = t.find("-->", p + 4); if (e == std::string::npos) fail("unterminated comment"); p = e + 3; }
            else if (starts("<?")) { size_t e = t.find("?>", p + 2); if (e == std::string::npos) fail("unterminated declaration"); p = e + 2; }
            else if (starts("<!")) { size_t e = t.find('>', p); if (e == std::string::npos) fail("unterminated <!"); p = e + 1; }
            else break;
        }
    }
    std::string parseName() {
        size_t b = p;
        while (p < t.size() && (std::isalnum((unsigned char)t[p]) || t[p] == '_' || t[p] == '-' || t[p] == ':' || t[p] == '.')) p++;
        if (b == p) fail("expected a name");
        return t.substr(b, p - b);
    }
    static std::string unescape(const std::string& s) {
        std::string o;
        for (size_t i = 0; i < s.size(); i++) {
            if (s[i] != '&') { o += s[i]; continue; }
            static const std::pair<const char*, char> ents[] = {{"&amp;", '&'}, {"&lt;", '<'}, {"&gt;", '>'}, {"&quot;", '"'}, {"&apos;", '\''}};
            bool hit = false;
            for (auto& e : ents) {
                size_t n = std::char_traits<char>::length(e.first);
                if (s.compare(i, n, e.first) == 0) { o += e.second; i += n - 1; hit = true; break; }
            }
            if (!hit) o += '&';
        }
        return o;
    }
    std::shared_ptr<XmlNode::Data> parseElement() {
        p++;   // '<'
        auto node = std::make_shared<XmlNode::Data>();
        node->name = parseName();
        while (true) {
            skipWs();
            if (p >= t.size()) fail("unterminated tag");
            if (starts("/>")) { p += 2; return node; }
            if (t[p] == '>') { p++; break; }
            std::string an = parseName();
            skipWs();
            if (p >= t.size() || t[p] != '=') fail("expected '='");
            p++;
            skipWs();
            if (p >= t.size() || (t[p] != '"' && t[p] != '\'')) fail("expected a quoted value");
            char q = t[p++];
            size_t e = t.find(q, p);
            if (e == std::string::npos) fail("unterminated attribute value");
            node->attrs.emplace_back(an, unescape(t.substr(p, e - p)));
            p = e + 1;
        }
        while (true) {   // content
            size_t lt = t.find('<', p);
            if (lt == std::string::npos) fail("unterminated element <" + node->name + ">");
            p = lt;      // character data is ignored: scene.xml carries everything in attributes
            if (starts("<!--") || starts("<?") || starts("<!")) { skipMisc(); continue; }
            if (starts("</")) {
                p += 2;
                std::string cn = parseName();
                if (cn != node->name) fail("mismatched </" + cn + ">");
                skipWs();
                if (p >= t.size() || t[p] != '>') fail("expected '>'");
                p++;
                return node;
            }
            node->children.push_back(parseElement());
        }
    }
};

XmlNode XmlNode::parseString(const std::string& text, std::string* error) {
    XmlParser parser(text);
    auto d = parser.parseDocument(error);
    return d ? XmlNode(d) : XmlNode();
}

XmlNode XmlNode::parseFile(const std::string& path, std::string* error) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { if (error) *error = "cannot open " + path; return XmlNode(); }
    std::stringstream ss;
    ss << f.rdbuf();
    return parseString(ss.str(), error);
}

}  // namespace zillum
