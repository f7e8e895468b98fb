#pragma once
#include "../stub_core.hh"
