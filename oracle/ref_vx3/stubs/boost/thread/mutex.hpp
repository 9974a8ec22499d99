// TEST INFRASTRUCTURE — stand-in for boost::mutex (src/old/VX3_MemoryCleaner.h).
#pragma once
#include <mutex>
namespace boost { typedef std::mutex mutex; }
