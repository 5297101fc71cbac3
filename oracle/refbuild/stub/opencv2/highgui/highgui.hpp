/* STUB OpenCV — TEST INFRASTRUCTURE ONLY (see opencv2/core/core.hpp). Nothing of highgui is used on the path. */
#include "opencv2/core/core.hpp"
