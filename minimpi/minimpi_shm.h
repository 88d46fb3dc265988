/* Layout of the shared segment of the minimpi multi-process back-end; shared by
 * minimpi_shm.c (the ranks) and minimpirun.c (the launcher that creates it). */
#ifndef MINIMPI_SHM_H
#define MINIMPI_SHM_H
#include <stdatomic.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define MINIMPI_MAGIC 0x6d696e696d706901ULL

typedef struct {
   uint64_t magic;
   int size;                 /* ranks */
   int pad;
   uint64_t ring_bytes;      /* capacity of one ring: power of two, >= 4096 */
   uint64_t rings_off;       /* offset of ring (0,0) */
   _Atomic int abort_flag;   /* a rank died or called MPI_Abort: everybody leaves */
} minimpi_seg;

static inline size_t minimpi_seg_bytes(int size, size_t ring_bytes)
{
   return 4096 + (size_t) size*(size_t) size*(128 + ring_bytes);
}

static inline void minimpi_seg_init(minimpi_seg *s, int size, size_t ring_bytes)
{
   memset(s, 0, 4096);
   s->magic = MINIMPI_MAGIC;
   s->size = size;
   s->ring_bytes = ring_bytes;
   s->rings_off = 4096;
}
#endif
