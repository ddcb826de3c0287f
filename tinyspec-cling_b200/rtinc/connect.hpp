// rtinc/connect.hpp — JACK port patching (reference: rtinc/connect.hpp:1-5) is audio-server wiring, not the
// data path: offline these are no-ops with the same signatures, so hacks that call them compile and run.
#pragma once
#include <string>
#define CONNECT_MIN 0
#define CONNECT_MAX 1
#define CONNECT_ALL 2
inline void connect(std::string, std::string, std::string = "", std::string = "", int = CONNECT_MAX) {}
inline void disconnect(std::string, std::string, std::string = "", std::string = "", int = CONNECT_MAX) {}
