/* Oracle shim (test infrastructure, not product code).
 * Host-x86 stand-in for the reference's OVI40 UI board header
 * (reference: mchf-eclipse/hardware/board_configs/UHSDR_UI_ovi40_config.h:28-78).
 * It keeps only the switches the RX/TX block path reads (two-channel audio,
 * two codecs, no special memory sections) and supplies dummy MCU types so the
 * reference's hot-path sources compile unmodified with gcc on x86-64.
 * Placed FIRST on the include path so it shadows the real board header. */
#ifndef ORACLE_SHIM_UI_OVI40_CONFIG_H
#define ORACLE_SHIM_UI_OVI40_CONFIG_H

#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>

/* newlib declares pow10f in <math.h>; glibc >= 2.27 no longer does.  Without a prototype the
 * reference's calls (audio_driver.c:909,937, audio_agc.c:229,285,318,331, audio_management.c:18)
 * would be compiled with an implicit int(...) signature.  ref_stubs.c defines it as exp10f. */
float pow10f(float x);

#define __MCHF_SPECIALMEM
#define __UHSDR_DMAMEM
#define USE_TWO_CHANNEL_AUDIO
#define CODEC_NUM 2
#define TRX_NAME "oracle-x86"
#define TRX_ID "orcl"

#ifndef __packed
#define __packed __attribute__((packed))
#endif
#ifndef __IO
#define __IO volatile
#endif
#ifndef __weak
#define __weak __attribute__((weak))
#endif

typedef struct { volatile uint32_t BSRR, ODR, IDR; } GPIO_TypeDef;
typedef struct { int dummy; } I2C_HandleTypeDef;
typedef struct { int dummy; } SPI_HandleTypeDef;
typedef struct { int dummy; } RTC_HandleTypeDef;
typedef struct { int dummy; } DAC_HandleTypeDef;
typedef struct { int dummy; } TIM_HandleTypeDef;
typedef struct { int dummy; } UART_HandleTypeDef;
typedef enum { HAL_OK = 0, HAL_ERROR = 1, HAL_BUSY = 2, HAL_TIMEOUT = 3 } HAL_StatusTypeDef;

typedef struct { volatile uint32_t ICSR; } oracle_scb_t;
extern oracle_scb_t shim_scb;
#define SCB (&shim_scb)
#define SCB_ICSR_PENDSVSET_Msk (1u << 28)

#endif
