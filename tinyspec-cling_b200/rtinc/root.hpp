// rtinc/root.hpp — everything a hack sees, as in the reference's rtinc/root.hpp + rtlib/root.hpp
// (`using namespace std`, RATE / CLIENT_NAME, buffers, control, connect, osc, fft, windows, phase vocoder).
#pragma once
#include <cmath>
#include <complex>
#include <cstring>
#include <iostream>
#include <vector>
using namespace std;
#include "buffer.hpp"
#include "control.hpp"
#include "connect.hpp"
#include "osc.hpp"
#include "fft.hpp"
#include "windowfn.hpp"
#include "pv.hpp"
#include "driver.hpp"
