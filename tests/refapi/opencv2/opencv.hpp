// stand-in for the third-party header of the same name (absent from this image): see host/compat_types.h
#pragma once
#include "../../../triangulation-in-deformable-scenes_b200/host/compat_types.h"
