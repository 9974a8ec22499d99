#pragma once
#include "curand.h"
