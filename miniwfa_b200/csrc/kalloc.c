/*
 * kalloc.c -- arena allocator behind the kalloc API (reference kalloc.c:38-224 for the
 * behaviour each entry point must have; the implementation here is independent).
 *
 * Design: an arena owns a list of "cores" obtained from its parent arena (or from libc
 * when it has none).  Inside a core, memory is cut into blocks that carry a 16-byte
 * header; free blocks form one address-ordered singly linked list per arena, searched
 * first-fit, split from the front and coalesced with both neighbours on release.
 * Payloads are 16-byte aligned.  As in the reference: kmalloc(km,0) == NULL, a NULL arena
 * forwards to malloc/calloc/realloc/free, krealloc never shrinks, krelocate moves a
 * block to a fresh allocation (so a result can outlive the scratch around it), and
 * km_destroy hands every core back to the parent.  No locking: one arena per thread.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "kalloc.h"

#define KA_UNIT       16u        /* accounting granule; min_core_size is counted in these (kalloc.c:41-42) */
#define KA_DEF_UNITS  0x80000u   /* default core: 8 MiB */
#define KA_MIN_SPLIT  32u        /* do not leave free fragments smaller than this */

typedef struct ka_blk {
	size_t size;            /* bytes of the whole block, header included; multiple of 16 */
	struct ka_blk *next;    /* next free block by address (free blocks only) */
} ka_blk_t;

typedef struct ka_core {
	struct ka_core *next;
	size_t size;            /* bytes of the whole core, this header included */
} ka_core_t;

typedef struct {
	void *parent;
	size_t min_core_units;
	ka_core_t *cores;
	ka_blk_t *free_head;
} ka_arena_t;

static void ka_die(const char *msg)
{
	fprintf(stderr, "[kalloc] %s\n", msg);
	abort();
}

void *km_init2(void *km_par, size_t min_core_size)
{
	ka_arena_t *a = (ka_arena_t*)kcalloc(km_par, 1, sizeof(ka_arena_t));
	if (a == 0) ka_die("cannot allocate an arena");
	a->parent = km_par;
	if (min_core_size > 0) a->min_core_units = min_core_size;
	else if (km_par) a->min_core_units = ((ka_arena_t*)km_par)->min_core_units - 2; /* child cores fit in a parent core */
	else a->min_core_units = KA_DEF_UNITS;
	return a;
}

void *km_init(void) { return km_init2(0, 0); }

void km_destroy(void *km)
{
	ka_arena_t *a = (ka_arena_t*)km;
	ka_core_t *c, *nx;
	void *par;
	if (a == 0) return;
	par = a->parent;
	for (c = a->cores; c; c = nx) {
		nx = c->next;
		kfree(par, c);
	}
	kfree(par, a);
}

/* put a block on the address-ordered free list, merging with adjacent free blocks */
static void ka_release(ka_arena_t *a, ka_blk_t *b)
{
	ka_blk_t *prev = 0, *cur = a->free_head;
	while (cur && cur < b) prev = cur, cur = cur->next;
	if (cur == b || (prev && (char*)prev + prev->size > (char*)b) || (cur && (char*)b + b->size > (char*)cur))
		ka_die("kfree: block overlaps free memory (double free or corruption)");
	if (cur && (char*)b + b->size == (char*)cur) { /* merge with the successor */
		b->size += cur->size;
		b->next = cur->next;
	} else b->next = cur;
	if (prev && (char*)prev + prev->size == (char*)b) { /* merge with the predecessor */
		prev->size += b->size;
		prev->next = b->next;
	} else if (prev) prev->next = b;
	else a->free_head = b;
}

static void ka_grow(ka_arena_t *a, size_t need_bytes)
{
	size_t unit_bytes = a->min_core_units * KA_UNIT, bytes;
	ka_core_t *c;
	ka_blk_t *b;
	if (unit_bytes < 4096) unit_bytes = 4096;
	bytes = need_bytes + sizeof(ka_core_t);
	bytes = (bytes + unit_bytes - 1) / unit_bytes * unit_bytes;
	c = (ka_core_t*)kmalloc(a->parent, bytes);
	if (c == 0) ka_die("out of memory while growing an arena");
	c->next = a->cores, c->size = bytes, a->cores = c;
	b = (ka_blk_t*)(c + 1); /* the core header keeps blocks of neighbouring cores from merging */
	b->size = (bytes - sizeof(ka_core_t)) & ~(size_t)(KA_UNIT - 1);
	ka_release(a, b);
}

void *kmalloc(void *km, size_t n_bytes)
{
	ka_arena_t *a = (ka_arena_t*)km;
	size_t need;
	int pass;
	if (n_bytes == 0) return 0;
	if (a == 0) return malloc(n_bytes);
	need = (n_bytes + sizeof(ka_blk_t) + (KA_UNIT - 1)) & ~(size_t)(KA_UNIT - 1);
	for (pass = 0; pass < 2; ++pass) {
		ka_blk_t *prev = 0, *cur;
		for (cur = a->free_head; cur; prev = cur, cur = cur->next) {
			if (cur->size < need) continue;
			if (cur->size - need >= KA_MIN_SPLIT) { /* hand out the front, keep the tail free */
				ka_blk_t *rest = (ka_blk_t*)((char*)cur + need);
				rest->size = cur->size - need;
				rest->next = cur->next;
				cur->size = need;
				if (prev) prev->next = rest; else a->free_head = rest;
			} else {
				if (prev) prev->next = cur->next; else a->free_head = cur->next;
			}
			cur->next = 0;
			return (char*)cur + sizeof(ka_blk_t);
		}
		ka_grow(a, need);
	}
	ka_die("kmalloc: no fit after growing");
	return 0;
}

void kfree(void *km, void *p)
{
	if (p == 0) return;
	if (km == 0) { free(p); return; }
	ka_release((ka_arena_t*)km, (ka_blk_t*)((char*)p - sizeof(ka_blk_t)));
}

void *kcalloc(void *km, size_t count, size_t size)
{
	void *p;
	if (count == 0 || size == 0) return 0;
	if (km == 0) return calloc(count, size);
	p = kmalloc(km, count * size);
	memset(p, 0, count * size);
	return p;
}

void *krealloc(void *km, void *p, size_t n_bytes)
{
	size_t have;
	void *q;
	if (n_bytes == 0) { kfree(km, p); return 0; }
	if (km == 0) return realloc(p, n_bytes);
	if (p == 0) return kmalloc(km, n_bytes);
	have = ((ka_blk_t*)((char*)p - sizeof(ka_blk_t)))->size - sizeof(ka_blk_t);
	if (have >= n_bytes) return p; /* never shrinks, like the reference (kalloc.c:166) */
	q = kmalloc(km, n_bytes);
	memcpy(q, p, have);
	kfree(km, p);
	return q;
}

void *krelocate(void *km, void *p, size_t n_bytes)
{
	void *q;
	if (km == 0) return p;
	q = kmalloc(km, n_bytes);
	if (n_bytes) memcpy(q, p, n_bytes);
	kfree(km, p);
	return q;
}

void km_stat(const void *km, km_stat_t *s)
{
	const ka_arena_t *a = (const ka_arena_t*)km;
	const ka_blk_t *b;
	const ka_core_t *c;
	memset(s, 0, sizeof(*s));
	if (a == 0) return;
	for (b = a->free_head; b; b = b->next) s->available += b->size, ++s->n_blocks;
	for (c = a->cores; c; c = c->next) {
		++s->n_cores, s->capacity += c->size;
		if (c->size > s->largest) s->largest = c->size;
	}
}

void km_stat_print(const void *km)
{
	km_stat_t st;
	km_stat(km, &st);
	fprintf(stderr, "[km_stat] cap=%ld, avail=%ld, largest=%ld, n_core=%ld, n_block=%ld\n",
			(long)st.capacity, (long)st.available, (long)st.largest, (long)st.n_cores, (long)st.n_blocks);
}
