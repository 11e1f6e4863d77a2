/* TEST INFRASTRUCTURE: see opencv.hpp next to this file. */
#include "opencv.hpp"
