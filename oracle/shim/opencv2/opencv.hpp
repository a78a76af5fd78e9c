// TEST INFRASTRUCTURE ONLY: forwards to the OpenCV stand-in used to compile the reference sources.
#pragma once
#include "../cvshim.h"
