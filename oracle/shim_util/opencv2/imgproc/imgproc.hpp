#include "cvshim_util.hpp"
