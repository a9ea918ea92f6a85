/*
 * kalloc.h -- arena allocator with the API of lh3/miniwfa's kalloc
 * (reference kalloc.h:14-41).  A NULL arena handle means "use libc".
 * Not thread safe: one arena per thread, as in the reference.
 */
#ifndef _KALLOC_H_
#define _KALLOC_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	size_t capacity, available, n_blocks, n_cores, largest;
} km_stat_t;

void *kmalloc(void *km, size_t size);
void *krealloc(void *km, void *ptr, size_t size);
void *krelocate(void *km, void *ap, size_t n_bytes);
void *kcalloc(void *km, size_t count, size_t size);
void kfree(void *km, void *ptr);

void *km_init(void);
void *km_init2(void *km_par, size_t min_core_size);
void km_destroy(void *km);
void km_stat(const void *km, km_stat_t *s);
void km_stat_print(const void *km);

#ifdef __cplusplus
}
#endif

#define Kmalloc(km, type, cnt)       ((type*)kmalloc((km), (cnt) * sizeof(type)))
#define Kcalloc(km, type, cnt)       ((type*)kcalloc((km), (cnt), sizeof(type)))
#define Krealloc(km, type, ptr, cnt) ((type*)krealloc((km), (ptr), (cnt) * sizeof(type)))

#define KMALLOC(km, ptr, len)  ((ptr) = (__typeof__(ptr))kmalloc((km), (len) * sizeof(*(ptr))))
#define KCALLOC(km, ptr, len)  ((ptr) = (__typeof__(ptr))kcalloc((km), (len), sizeof(*(ptr))))
#define KREALLOC(km, ptr, len) ((ptr) = (__typeof__(ptr))krealloc((km), (ptr), (len) * sizeof(*(ptr))))

#define KEXPAND(km, a, m) do { \
		(m) = (m) >= 4? (m) + ((m)>>1) : 16; \
		KREALLOC((km), (a), (m)); \
	} while (0)

/*
 * Typed object pools on top of an arena (reference kalloc.h:43-80; unused by miniwfa itself, kept because klib users of
 * kalloc.h expect them).  KALLOC_POOL_INIT(name, type) defines kmp_name_t and kmp_init_name / kmp_destroy_name /
 * kmp_alloc_name / kmp_free_name: freed objects are parked on a stack and handed out again before the arena is asked;
 * a fresh object is zero-filled, a recycled one comes back as it was freed.
 */
#ifndef klib_unused
#if defined(__GNUC__) || defined(__clang__)
#define klib_unused __attribute__((__unused__))
#else
#define klib_unused
#endif
#endif

#define KALLOC_POOL_INIT2(SCOPE, name, kmptype_t) \
	typedef struct { size_t cnt, n, max; kmptype_t **buf; void *km; } kmp_##name##_t; \
	SCOPE kmp_##name##_t *kmp_init_##name(void *km) \
	{ \
		kmp_##name##_t *pool = (kmp_##name##_t*)kcalloc(km, 1, sizeof(kmp_##name##_t)); \
		pool->km = km; \
		return pool; \
	} \
	SCOPE void kmp_destroy_##name(kmp_##name##_t *pool) \
	{ \
		while (pool->n > 0) kfree(pool->km, pool->buf[--pool->n]); \
		kfree(pool->km, pool->buf); \
		kfree(pool->km, pool); \
	} \
	SCOPE kmptype_t *kmp_alloc_##name(kmp_##name##_t *pool) \
	{ \
		++pool->cnt; \
		return pool->n > 0 ? pool->buf[--pool->n] : (kmptype_t*)kcalloc(pool->km, 1, sizeof(kmptype_t)); \
	} \
	SCOPE void kmp_free_##name(kmp_##name##_t *pool, kmptype_t *obj) \
	{ \
		--pool->cnt; \
		if (pool->n == pool->max) KEXPAND(pool->km, pool->buf, pool->max); \
		pool->buf[pool->n++] = obj; \
	}

#define KALLOC_POOL_INIT(name, kmptype_t) KALLOC_POOL_INIT2(static inline klib_unused, name, kmptype_t)

#endif
