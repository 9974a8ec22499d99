// Minimal XML DOM reader/writer for the VXA / VXD / VXT / VXR files (the reference uses Boost.PropertyTree and a
// vendored TinyXML; neither is available / needed here).  Supports elements, attributes, text, CDATA, comments,
// processing instructions and the five predefined entities — everything the reference's files contain.
#pragma once
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace vx3 {

struct XNode {
    std::string name;
    std::string text; // concatenated character data (CDATA included) directly inside this element
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<std::unique_ptr<XNode>> kids;

    const XNode *child(const std::string &n) const;
    XNode *child(const std::string &n);
    std::vector<const XNode *> children(const std::string &n) const;
    const std::string *attr(const std::string &n) const;
    // dotted path lookup, first match per level (boost::property_tree::get_child semantics): "VXA.Simulator.X"
    const XNode *path(const std::string &dotted) const;
    std::unique_ptr<XNode> clone() const;
    // put_child semantics: create intermediate nodes, replace the first existing leaf or append
    void put(const std::string &dotted, std::unique_ptr<XNode> node);
};

// Parses a document; the returned node is a nameless root whose kids are the top-level elements.
std::unique_ptr<XNode> xml_parse(const std::string &src, std::string *err);
bool xml_read_file(const std::string &path, std::string *out);
std::string xml_escape(const std::string &s);

// text -> value helpers with the reference's "tag absent -> default" convention
std::string xml_trim(const std::string &s);
bool xml_get_double(const XNode *parent, const char *tag, double *out);
bool xml_get_int(const XNode *parent, const char *tag, int *out);
bool xml_get_bool(const XNode *parent, const char *tag, bool *out);

} // namespace vx3
