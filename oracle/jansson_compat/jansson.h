/*
 * Declarations-only stand-in for <jansson.h> (jansson 2.14 public API), TEST
 * INFRASTRUCTURE.  The image ships the runtime libjansson.so.4 but no header
 * (SURVEY D7); the reference's src/main.c needs exactly the subset below
 * (main.c:22-73,140-184).  Link with -l:libjansson.so.4.
 */
#ifndef NAVSLAM_JANSSON_COMPAT_H
#define NAVSLAM_JANSSON_COMPAT_H
#include <stddef.h>
#include <stdio.h>

typedef enum {
    JSON_OBJECT, JSON_ARRAY, JSON_STRING, JSON_INTEGER, JSON_REAL, JSON_TRUE, JSON_FALSE, JSON_NULL
} json_type;

typedef struct json_t {
    json_type type;
    volatile size_t refcount;
} json_t;

typedef long long json_int_t;

typedef struct json_error_t {
    int line;
    int column;
    int position;
    char source[80];
    char text[160];
} json_error_t;

#define json_typeof(json) ((json)->type)
#define json_is_object(json) ((json) && json_typeof(json) == JSON_OBJECT)
#define json_is_array(json) ((json) && json_typeof(json) == JSON_ARRAY)
#define json_is_integer(json) ((json) && json_typeof(json) == JSON_INTEGER)
#define json_is_real(json) ((json) && json_typeof(json) == JSON_REAL)

json_t *json_loadf(FILE *input, size_t flags, json_error_t *error);
size_t json_array_size(const json_t *array);
json_t *json_array_get(const json_t *array, size_t index);
json_t *json_object_get(const json_t *object, const char *key);
json_int_t json_integer_value(const json_t *integer);
double json_real_value(const json_t *real);
void json_delete(json_t *json);

static inline void json_decref(json_t *json) {
    if (json && json->refcount != (size_t)-1 &&
        __atomic_sub_fetch(&json->refcount, 1, __ATOMIC_RELEASE) == 0)
        json_delete(json);
}

#define json_array_foreach(array, index, value)                                   \
    for (index = 0; index < json_array_size(array) && (value = json_array_get(array, index)); \
         index++)

#endif
