#include "cvshim_dev.hpp"
