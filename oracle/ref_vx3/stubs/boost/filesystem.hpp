// TEST INFRASTRUCTURE — minimal stand-in so the reference's src/Utils/VX3.cuh parses without Boost (absent in this image).
// Only what VX3.cuh's two inline helpers name; the physics sources never use it.
#pragma once
#include <string>
namespace boost { namespace filesystem {
class path {
    std::string s_;
public:
    path() {}
    path(const std::string &s) : s_(s) {}
    path(const char *s) : s_(s) {}
    path filename() const { size_t p = s_.find_last_of('/'); return path(p == std::string::npos ? s_ : s_.substr(p + 1)); }
    path extension() const { size_t p = s_.find_last_of('.'); return path(p == std::string::npos ? std::string() : s_.substr(p)); }
    std::string string() const { return s_; }
    const char *c_str() const { return s_.c_str(); }
};
} }
