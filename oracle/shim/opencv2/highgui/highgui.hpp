// TEST INFRASTRUCTURE ONLY: forwards to the OpenCV stand-in used to compile the reference extractor.
#pragma once
#include "../../cvshim.h"
