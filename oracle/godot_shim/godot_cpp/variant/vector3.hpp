#pragma once
#include "../../godot_shim.hpp"
