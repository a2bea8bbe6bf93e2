/* oracle/shim/sonLib.h -- TEST INFRASTRUCTURE ONLY.  Minimal stand-in for the sonLib
 * containers used by the reference's marker path (stList: growable pointer array with an
 * element destructor; stHash: string-keyed map).  Written from the call sites' semantics
 * (construct3/append/get/length/sort/copy/destruct, search/insert/iterator/getKeys). */
#ifndef ORACLE_SHIM_SONLIB_H
#define ORACLE_SHIM_SONLIB_H
#include <stdint.h>
#include <stdbool.h>
#include <stdlib.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct stList stList;
typedef struct stHash stHash;
typedef struct stHashIterator stHashIterator;

stList *stList_construct3(int64_t size, void (*destructElement)(void *));
void stList_destruct(stList *list);
void stList_append(stList *list, void *item);
void *stList_get(stList *list, int64_t index);
int64_t stList_length(stList *list);
void stList_sort(stList *list, int (*cmpFn)(const void *a, const void *b));
stList *stList_copy(stList *list, void (*destructItem)(void *));

uint64_t stHash_stringKey(const void *k);
int stHash_stringEqualKey(const void *key1, const void *key2);
stHash *stHash_construct3(uint64_t (*hashKey)(const void *), int (*hashEqualsKey)(const void *, const void *),
                          void (*destructKeys)(void *), void (*destructValues)(void *));
void stHash_destruct(stHash *hash);
void stHash_insert(stHash *hash, void *key, void *value);
void *stHash_search(stHash *hash, void *key);
stHashIterator *stHash_getIterator(stHash *hash);
void *stHash_getNext(stHashIterator *iterator);
void stHash_destructIterator(stHashIterator *iterator);
stList *stHash_getKeys(stHash *hash);
#ifdef __cplusplus
}
#endif
#endif
