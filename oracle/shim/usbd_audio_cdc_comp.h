/* Oracle shim: USB audio class constants the block path references
 * (reference asserts USBD_AUDIO_FREQ == 48000, audio_driver.c:2632-2635). */
#ifndef ORACLE_SHIM_USBD_AUDIO_CDC_COMP_H
#define ORACLE_SHIM_USBD_AUDIO_CDC_COMP_H
#include <stdint.h>
#define USBD_AUDIO_FREQ 48000
typedef struct { int dummy; } USBD_AUDIO_ItfTypeDef;
#endif
