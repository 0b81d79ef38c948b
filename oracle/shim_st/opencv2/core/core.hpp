#include "cvshim_st.hpp"
