// TEST INFRASTRUCTURE — stand-in for the one Boost string helper src/Utils/VX3.cuh names.
#pragma once
#include <cctype>
#include <string>
namespace boost {
inline void to_upper(std::string &s) { for (auto &c : s) c = (char)std::toupper((unsigned char)c); }
}
