// rtinc/osc.hpp — OSC messaging (reference: rtinc/osc.hpp:27-45) is control-plane I/O outside the hot path:
// offline, nothing is sent and nothing is ever received.  Same names and call shapes, so hacks compile.
#pragma once
#include <memory>
#include <string>
#include <vector>
struct RecvMsg {};
template <typename... T>
void osc_send(const std::string &, int, double, const std::string &, T...) {}
inline std::vector<std::shared_ptr<RecvMsg>> osc_recv(int, double, const std::string &) { return {}; }
inline bool osc_get(RecvMsg &) { return true; }
template <typename T1, typename... T>
bool osc_get(RecvMsg &, T1 &, T &...) { return false; }
