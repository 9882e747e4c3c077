// oracle/shim: parameters.h includes OpenCV only for its YAML loader, which is not compiled here
#pragma once
