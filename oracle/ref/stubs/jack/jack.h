/* Declaration-only stand-in for <jack/jack.h> (JACK public API subset).
 * TEST INFRASTRUCTURE ONLY: the 13 functions the unmodified reference uses
 * (tinyspec.cpp:205-258, rtlib/control.hpp:54-77, rtlib/connect.hpp); the offline
 * fake driver in ../fake_jack.cpp defines them. */
#ifndef TSREF_JACK_STUB_H
#define TSREF_JACK_STUB_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef uint32_t jack_nframes_t;
typedef struct _jack_client jack_client_t;
typedef struct _jack_port jack_port_t;
typedef int jack_status_t;
typedef int jack_options_t;
enum { JackNullOption = 0 };
enum JackPortFlags { JackPortIsInput = 0x1, JackPortIsOutput = 0x2, JackPortIsPhysical = 0x4 };
#define JACK_DEFAULT_AUDIO_TYPE "32 bit float mono audio"
typedef int (*JackProcessCallback)(jack_nframes_t nframes, void *arg);
typedef int (*JackSampleRateCallback)(jack_nframes_t nframes, void *arg);
jack_client_t *jack_client_open(const char *name, jack_options_t options, jack_status_t *status, ...);
char *jack_get_client_name(jack_client_t *client);
int jack_set_process_callback(jack_client_t *client, JackProcessCallback cb, void *arg);
int jack_set_sample_rate_callback(jack_client_t *client, JackSampleRateCallback cb, void *arg);
int jack_activate(jack_client_t *client);
int jack_deactivate(jack_client_t *client);
void *jack_port_get_buffer(jack_port_t *port, jack_nframes_t nframes);
jack_port_t *jack_port_register(jack_client_t *client, const char *port_name, const char *port_type,
                                unsigned long flags, unsigned long buffer_size);
int jack_port_unregister(jack_client_t *client, jack_port_t *port);
const char **jack_get_ports(jack_client_t *client, const char *port_name_pattern,
                            const char *type_name_pattern, unsigned long flags);
int jack_connect(jack_client_t *client, const char *source_port, const char *destination_port);
int jack_disconnect(jack_client_t *client, const char *source_port, const char *destination_port);
void jack_free(void *ptr);
#ifdef __cplusplus
}
#endif
#endif
