// ts_render — offline render driver: the main() a hack is compiled into, as the reference's tinyspec.cpp is
// (`./compile --no-cling hack.cpp` -> -DHACK=..., tinyspec.cpp:275-294), with the JACK client and the worker
// thread replaced by an offline scheduler on the device (rtinc/driver.hpp).
//
//   make -C tinyspec-cling_b200/driver HACK=/path/to/hack.cpp        # -> ts_render_<hack>
//   ./ts_render_<hack> --seconds 60 --streams 64 --precision f32 --out out.f32 --timing
#include "../rtinc/root.hpp"

#define TSB_STR_(x) #x
#define TSB_STR(x) TSB_STR_(x)

int main(int argc, char **argv) {
    tsbrt::Renderer renderer(tsbrt::parse_args(argc, argv));
    renderer.init();
#ifdef HACK
#include TSB_STR(HACK)
#else
#error "no hack specified: compile with -DHACK=path/to/hack.cpp"
#endif
    return renderer.run();
}
