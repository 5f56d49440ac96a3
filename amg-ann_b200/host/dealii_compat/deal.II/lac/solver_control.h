// deal.II-compat include shim: see ../../amgb_dealii_compat.hpp
#pragma once
#include "../../amgb_dealii_compat.hpp"
