// TEST INFRASTRUCTURE — src/old/VX3_MemoryCleaner.h includes it and uses nothing of it.
#pragma once
