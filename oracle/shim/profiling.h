/* Oracle shim (test infrastructure): host stand-in for the reference's misc/profiling.h:18-150.
 * The firmware's profiler reads the Cortex-M DWT cycle counter at 0xE0001004, which does not
 * exist on x86; the event table and the call names are kept, the timers are no-ops. */
#ifndef ORACLE_SHIM_PROFILING_H
#define ORACLE_SHIM_PROFILING_H
#include <stdint.h>
typedef enum {
    ProfileAudioInterrupt = 0, ProfileTP1, ProfileTP2, ProfileTP3, ProfileTP4, ProfileTP5, ProfileTP6,
    ProfileTP7, ProfileTP8, ProfileTP9, ProfileFreeDV, FreeDVTXUnderrun, EventProfileMax
} ProfiledEventNames;
typedef struct { uint32_t count; uint32_t start; uint32_t stop; uint64_t duration; } ProfilingTimedEvent;
typedef struct { ProfilingTimedEvent event[EventProfileMax]; } EventProfile_t;
extern EventProfile_t eventProfile;
static inline void profileEvent(const ProfiledEventNames pe) { if (pe < EventProfileMax) eventProfile.event[pe].count++; }
static inline void profileTimedEventInit(void) {}
static inline void profileTimedEventStart(const ProfiledEventNames pe) { (void)pe; }
static inline void profileTimedEventStop(const ProfiledEventNames pe) { if (pe < EventProfileMax) eventProfile.event[pe].count++; }
static inline void profileTimedEventReset(const ProfiledEventNames pe) { (void)pe; }
#endif
