/* node_api_min.h — hand-declared subset of Node's stable C N-API (node_api.h / js_native_api.h,
 * NAPI_VERSION 6) used by nd4b_napi.cc.  The image has no Node.js and therefore no node_api.h; these
 * declarations follow the published N-API signatures so that the addon can be COMPILE-checked here.
 * With a real Node toolchain, build with -DND4B_HAVE_NODE_API to include <node_api.h> instead. */
#ifndef ND4B_NODE_API_MIN_H
#define ND4B_NODE_API_MIN_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_callback_info__* napi_callback_info;
typedef enum { napi_ok = 0, napi_invalid_arg, napi_object_expected, napi_string_expected, napi_name_expected,
               napi_function_expected, napi_number_expected, napi_boolean_expected, napi_array_expected,
               napi_generic_failure, napi_pending_exception } napi_status;
typedef enum { napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array,
               napi_int32_array, napi_uint32_array, napi_float32_array, napi_float64_array,
               napi_bigint64_array, napi_biguint64_array } napi_typedarray_type;
typedef enum { napi_default = 0 } napi_property_attributes;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void* finalize_data, void* finalize_hint);
typedef struct {
  const char* utf8name; napi_value name; napi_callback method; napi_callback getter; napi_callback setter;
  napi_value value; napi_property_attributes attributes; void* data;
} napi_property_descriptor;
typedef napi_value (*napi_addon_register_func)(napi_env env, napi_value exports);
typedef struct napi_module {
  int nm_version; unsigned int nm_flags; const char* nm_filename; napi_addon_register_func nm_register_func;
  const char* nm_modname; void* nm_priv; void* reserved[4];
} napi_module;

napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t* argc, napi_value* argv,
                             napi_value* this_arg, void** data);
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type* type, size_t* length,
                                     void** data, napi_value* arraybuffer, size_t* byte_offset);
napi_status napi_get_array_length(napi_env env, napi_value value, uint32_t* result);
napi_status napi_get_element(napi_env env, napi_value object, uint32_t index, napi_value* result);
napi_status napi_get_value_int32(napi_env env, napi_value value, int32_t* result);
napi_status napi_get_value_int64(napi_env env, napi_value value, int64_t* result);
napi_status napi_get_value_double(napi_env env, napi_value value, double* result);
napi_status napi_create_int32(napi_env env, int32_t value, napi_value* result);
napi_status napi_create_double(napi_env env, double value, napi_value* result);
napi_status napi_create_string_utf8(napi_env env, const char* str, size_t length, napi_value* result);
napi_status napi_create_object(napi_env env, napi_value* result);
napi_status napi_set_named_property(napi_env env, napi_value object, const char* utf8name, napi_value value);
napi_status napi_get_undefined(napi_env env, napi_value* result);
napi_status napi_throw_error(napi_env env, const char* code, const char* msg);
napi_status napi_define_properties(napi_env env, napi_value object, size_t property_count,
                                   const napi_property_descriptor* properties);
napi_status napi_create_external_arraybuffer(napi_env env, void* external_data, size_t byte_length,
                                             napi_finalize finalize_cb, void* finalize_hint, napi_value* result);
napi_status napi_create_typedarray(napi_env env, napi_typedarray_type type, size_t length, napi_value arraybuffer,
                                   size_t byte_offset, napi_value* result);
void napi_module_register(napi_module* mod);
#ifdef __cplusplus
}
#endif
#define NAPI_AUTO_LENGTH SIZE_MAX
#endif
