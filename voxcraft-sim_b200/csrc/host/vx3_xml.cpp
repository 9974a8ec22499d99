#include "vx3_xml.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace vx3 {

const XNode *XNode::child(const std::string &n) const {
    for (auto &k : kids)
        if (k->name == n) return k.get();
    return nullptr;
}
XNode *XNode::child(const std::string &n) { return const_cast<XNode *>(static_cast<const XNode *>(this)->child(n)); }
std::vector<const XNode *> XNode::children(const std::string &n) const {
    std::vector<const XNode *> out;
    for (auto &k : kids)
        if (k->name == n) out.push_back(k.get());
    return out;
}
const std::string *XNode::attr(const std::string &n) const {
    for (auto &a : attrs)
        if (a.first == n) return &a.second;
    return nullptr;
}
const XNode *XNode::path(const std::string &dotted) const {
    const XNode *cur = this;
    size_t pos = 0;
    while (cur && pos <= dotted.size()) {
        size_t dot = dotted.find('.', pos);
        std::string part = dotted.substr(pos, dot == std::string::npos ? std::string::npos : dot - pos);
        if (!part.empty()) cur = cur->child(part);
        if (dot == std::string::npos) break;
        pos = dot + 1;
    }
    return cur;
}
std::unique_ptr<XNode> XNode::clone() const {
    std::unique_ptr<XNode> n(new XNode());
    n->name = name;
    n->text = text;
    n->attrs = attrs;
    for (auto &k : kids) n->kids.push_back(k->clone());
    return n;
}
void XNode::put(const std::string &dotted, std::unique_ptr<XNode> node) {
    XNode *cur = this;
    size_t pos = 0;
    std::vector<std::string> parts;
    while (true) {
        size_t dot = dotted.find('.', pos);
        parts.push_back(dotted.substr(pos, dot == std::string::npos ? std::string::npos : dot - pos));
        if (dot == std::string::npos) break;
        pos = dot + 1;
    }
    for (size_t i = 0; i + 1 < parts.size(); i++) {
        XNode *nx = cur->child(parts[i]);
        if (!nx) {
            cur->kids.emplace_back(new XNode());
            nx = cur->kids.back().get();
            nx->name = parts[i];
        }
        cur = nx;
    }
    node->name = parts.back();
    for (auto &k : cur->kids)
        if (k->name == parts.back()) {
            k = std::move(node);
            return;
        }
    cur->kids.push_back(std::move(node));
}

namespace {
struct Parser {
    const std::string &s;
    size_t i = 0;
    std::string err;
    explicit Parser(const std::string &src) : s(src) {}
    bool starts(const char *lit) const { return s.compare(i, strlen(lit), lit) == 0; }
    void skip_ws() {
        while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) i++;
    }
    static std::string decode(const std::string &t) {
        std::string o;
        o.reserve(t.size());
        for (size_t k = 0; k < t.size(); k++) {
            if (t[k] != '&') { o += t[k]; continue; }
            size_t e = t.find(';', k);
            if (e == std::string::npos) { o += t[k]; continue; }
            std::string ent = t.substr(k + 1, e - k - 1);
            if (ent == "lt") o += '<';
            else if (ent == "gt") o += '>';
            else if (ent == "amp") o += '&';
            else if (ent == "quot") o += '"';
            else if (ent == "apos") o += '\'';
            else if (!ent.empty() && ent[0] == '#') o += (char)strtol(ent.c_str() + (ent[1] == 'x' ? 2 : 1), nullptr, ent[1] == 'x' ? 16 : 10);
            else { o += t.substr(k, e - k + 1); }
            k = e;
        }
        return o;
    }
    bool skip_misc() { // comments, PIs, doctype; returns false on malformed input
        while (true) {
            skip_ws();
            if (starts("<!--")) {
                size_t e = s.find("-->", i);
                if (e == std::string::npos) { err = "unterminated comment"; return false; }
                i = e + 3;
            } else if (starts("<?")) {
                size_t e = s.find("?>", i);
                if (e == std::string::npos) { err = "unterminated processing instruction"; return false; }
                i = e + 2;
            } else if (starts("<!DOCTYPE")) {
                size_t e = s.find('>', i);
                if (e == std::string::npos) { err = "unterminated doctype"; return false; }
                i = e + 1;
            } else
                return true;
        }
    }
    std::string name() {
        size_t b = i;
        while (i < s.size() && !strchr(" \t\r\n/>=", s[i])) i++;
        return s.substr(b, i - b);
    }
    bool element(XNode &n) {
        // at '<'
        i++;
        n.name = name();
        if (n.name.empty()) { err = "empty element name"; return false; }
        while (true) {
            skip_ws();
            if (i >= s.size()) { err = "unexpected end inside tag " + n.name; return false; }
            if (s[i] == '/') {
                if (i + 1 < s.size() && s[i + 1] == '>') { i += 2; return true; }
                err = "malformed tag " + n.name;
                return false;
            }
            if (s[i] == '>') { i++; break; }
            std::string an = name();
            skip_ws();
            if (i >= s.size() || s[i] != '=') { err = "attribute without value in " + n.name; return false; }
            i++;
            skip_ws();
            if (i >= s.size() || (s[i] != '"' && s[i] != '\'')) { err = "unquoted attribute in " + n.name; return false; }
            char q = s[i++];
            size_t e = s.find(q, i);
            if (e == std::string::npos) { err = "unterminated attribute in " + n.name; return false; }
            n.attrs.emplace_back(an, decode(s.substr(i, e - i)));
            i = e + 1;
        }
        // content
        while (true) {
            if (i >= s.size()) { err = "unexpected end inside element " + n.name; return false; }
            if (s[i] != '<') {
                size_t e = s.find('<', i);
                if (e == std::string::npos) e = s.size();
                n.text += decode(s.substr(i, e - i));
                i = e;
                continue;
            }
            if (starts("<![CDATA[")) {
                size_t e = s.find("]]>", i);
                if (e == std::string::npos) { err = "unterminated CDATA"; return false; }
                n.text += s.substr(i + 9, e - i - 9);
                i = e + 3;
            } else if (starts("<!--")) {
                size_t e = s.find("-->", i);
                if (e == std::string::npos) { err = "unterminated comment"; return false; }
                i = e + 3;
            } else if (starts("<?")) {
                size_t e = s.find("?>", i);
                if (e == std::string::npos) { err = "unterminated processing instruction"; return false; }
                i = e + 2;
            } else if (starts("</")) {
                i += 2;
                std::string cn = name();
                skip_ws();
                if (i >= s.size() || s[i] != '>') { err = "malformed closing tag " + cn; return false; }
                i++;
                if (cn != n.name) { err = "mismatched closing tag: <" + n.name + "> closed by </" + cn + ">"; return false; }
                return true;
            } else {
                n.kids.emplace_back(new XNode());
                if (!element(*n.kids.back())) return false;
            }
        }
    }
};
} // namespace

std::unique_ptr<XNode> xml_parse(const std::string &src, std::string *err) {
    Parser p(src);
    std::unique_ptr<XNode> root(new XNode());
    while (true) {
        if (!p.skip_misc()) break;
        if (p.i >= src.size()) break;
        if (src[p.i] != '<') { p.err = "text outside of the root element"; break; }
        root->kids.emplace_back(new XNode());
        if (!p.element(*root->kids.back())) break;
    }
    if (!p.err.empty()) {
        if (err) *err = "XML parse error: " + p.err;
        return nullptr;
    }
    if (root->kids.empty()) {
        if (err) *err = "XML parse error: no root element";
        return nullptr;
    }
    return root;
}

bool xml_read_file(const std::string &path, std::string *out) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    out->clear();
    char buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) out->append(buf, n);
    fclose(f);
    return true;
}

std::string xml_escape(const std::string &s) {
    std::string o;
    for (char c : s) {
        if (c == '<') o += "&lt;";
        else if (c == '>') o += "&gt;";
        else if (c == '&') o += "&amp;";
        else o += c;
    }
    return o;
}

std::string xml_trim(const std::string &s) {
    size_t b = 0, e = s.size();
    while (b < e && strchr(" \t\r\n", s[b])) b++;
    while (e > b && strchr(" \t\r\n", s[e - 1])) e--;
    return s.substr(b, e - b);
}
bool xml_get_double(const XNode *parent, const char *tag, double *out) {
    const XNode *n = parent ? parent->child(tag) : nullptr;
    if (!n) return false;
    *out = atof(n->text.c_str()); // CXML_Rip::FindLoadElement semantics
    return true;
}
bool xml_get_int(const XNode *parent, const char *tag, int *out) {
    const XNode *n = parent ? parent->child(tag) : nullptr;
    if (!n) return false;
    *out = atoi(n->text.c_str());
    return true;
}
bool xml_get_bool(const XNode *parent, const char *tag, bool *out) {
    const XNode *n = parent ? parent->child(tag) : nullptr;
    if (!n) return false;
    std::string t = xml_trim(n->text);
    if (t == "true" || t == "True" || t == "TRUE") *out = true;
    else if (t == "false" || t == "False" || t == "FALSE") *out = false;
    else *out = atoi(t.c_str()) != 0;
    return true;
}

} // namespace vx3
