/*
 * dfgpu.h — C ABI of libdfgpu.so: the B200-native replacement for the DeepFlow
 * server ingester's per-record hot path (decode -> SmartEncoding -> flow_tag /
 * flow_metrics reduce -> ClickHouse column pack).
 *
 * Every entry point is what a cgo shim sitting in the slot of
 *   server/ingester/flow_log/decoder/decoder.go:106-200   (Decoder / Run)
 *   server/ingester/flow_metrics/unmarshaller/unmarshaller.go:95-282
 * would bind.  Plain C types only: pointers + sizes, no C++/torch types.
 * All pointers are caller-owned unless stated; the library never aborts, it
 * returns a negative errno-style code.
 *
 * Input side  (what the droplet queue hands a decoder, decoder.go:155-176):
 *   a batch of <=1024 *receiver.RecvBuffer (receiver.go:69-78): the shim
 *   concatenates Buffer[Begin:End] of every item into one staging buffer and
 *   describes each with a dfgpu_frame {offset,len,vtap,org,team}.
 * Output side (what ckwriter consumes, ckwriter.go:221-227, ckdb.go:33-36):
 *   per table one column block whose columns are laid out exactly like the
 *   ch-go v0.69 proto.Col* containers the reference appends to
 *   (l4_column_block.go:822-833, l7_column_block.go:167-169,376-422,
 *    flow_tag_column_block.go:50-101, tag_column_block.go:521-721).
 */
#ifndef DFGPU_H
#define DFGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes -------------------------------------------------------- */
#define DFGPU_OK 0
#define DFGPU_EINVAL (-22)
#define DFGPU_ENOMEM (-12)
#define DFGPU_ENODEV (-19)  /* no CUDA device / CUDA runtime error            */
#define DFGPU_ENOENT (-2)   /* unknown org / table / missing dictionary       */
#define DFGPU_EMIXEDORG (-71) /* frames of one submit carry different org ids */

/* ---- message types (server/libs/datatype/droplet-message.go:36-62) ------ */
#define DFGPU_MSG_METRICS 3      /* MESSAGE_TYPE_METRICS      */
#define DFGPU_MSG_TAGGEDFLOW 4   /* MESSAGE_TYPE_TAGGEDFLOW   */
#define DFGPU_MSG_PROTOCOLLOG 5  /* MESSAGE_TYPE_PROTOCOLLOG  */

/* ---- output tables ------------------------------------------------------ */
#define DFGPU_TABLE_L4_FLOW_LOG 0        /* flow_log.l4_flow_log                 */
#define DFGPU_TABLE_L7_FLOW_LOG 1        /* flow_log.l7_flow_log                 */
#define DFGPU_TABLE_L7_CUSTOM_FIELD 2    /* flow_tag.l7_flow_log_custom_field     */
#define DFGPU_TABLE_L7_CUSTOM_FIELD_VALUE 3 /* flow_tag.l7_flow_log_custom_field_value */
#define DFGPU_TABLE_L7_APP_SERVICE 4     /* flow_tag.l7_flow_log_app_service      */
/* flow_metrics tables: id = DFGPU_TABLE_METRICS_BASE + MetricsTableID
 * (server/libs/flow-metrics/tag.go:471-485: NETWORK_1M=0 .. TRAFFIC_POLICY_1M=4,
 *  NETWORK_1S=5 .. APPLICATION_MAP_1S=8)                                         */
#define DFGPU_TABLE_METRICS_BASE 16
/* flow_tag.<application|application_map>_app_service rows of the unmarshaller's
 * AppServiceTagWriter (unmarshaller.go:309-327)                                  */
#define DFGPU_TABLE_METRICS_APP_SERVICE 25
/* output of dfgpu_metrics_reduce: same columns as the table that was reduced    */
#define DFGPU_TABLE_METRICS_REDUCED 31
#define DFGPU_TABLE_MAX 32

/* ---- ch-go container kinds (server/libs/ckdb/ckdb.go, *_column_block.go) -- */
enum dfgpu_ch_type {
  DFGPU_CH_UINT8 = 1,
  DFGPU_CH_UINT16 = 2,
  DFGPU_CH_UINT32 = 3,
  DFGPU_CH_UINT64 = 4,
  DFGPU_CH_INT32 = 5,
  DFGPU_CH_INT64 = 6,
  DFGPU_CH_FLOAT64 = 7,
  DFGPU_CH_IPV4 = 8,        /* proto.ColIPv4  = []uint32                     */
  DFGPU_CH_IPV6 = 9,        /* proto.ColIPv6  = [][16]byte                   */
  DFGPU_CH_DATETIME = 10,   /* proto.ColDateTime   = []uint32 seconds        */
  DFGPU_CH_DATETIME64 = 11, /* proto.ColDateTime64 = []int64, precision 6    */
  DFGPU_CH_STR = 12,        /* proto.ColStr {Buf []byte; Pos []Position}     */
  DFGPU_CH_LC_STR = 13,     /* proto.ColLowCardinality[string], dictionary coded:
                               keys in data (elem_size 1/2/4), dictionary in
                               dict_* in order of first appearance            */
  DFGPU_CH_ARR_UINT16 = 14, /* proto.ColArr[uint16]: arr_offsets + data      */
  DFGPU_CH_ARR_STR = 15,    /* proto.ColArr[string]: arr_offsets + ColStr    */
  DFGPU_CH_ARR_LC_STR = 16, /* Array(LowCardinality(String)); same layout as
                               ARR_STR, the shim appends into the LC column   */
  DFGPU_CH_ARR_FLOAT64 = 17,
  DFGPU_CH_NULLABLE_UINT8 = 18,  /* proto.ColNullable[T]{Nulls ColUInt8; Values} */
  DFGPU_CH_NULLABLE_INT32 = 19,
  DFGPU_CH_NULLABLE_INT64 = 20,
  DFGPU_CH_NULLABLE_UINT64 = 21,
  DFGPU_CH_LC_STR_ROWS = 22 /* proto.ColLowCardinality[string] handed over as the
                               per-row Values (ColStr layout: data=Buf, pos); the
                               dictionary is derived by ch-go's Prepare() exactly
                               as in the reference                              */
};

/* proto.Position of ch-go ColStr: byte range [start,end) inside Buf. */
typedef struct dfgpu_pos {
  int64_t start;
  int64_t end;
} dfgpu_pos;

/* One output column.  Memory is library-owned pinned host memory, valid until
 * dfgpu_batch_release().                                                      */
typedef struct dfgpu_col {
  const char* name;      /* ClickHouse column name (INSERT binds by name)      */
  int32_t ch_type;       /* enum dfgpu_ch_type                                 */
  int32_t elem_size;     /* bytes per element of data for fixed-width kinds    */
  uint64_t nrows;
  const void* data;      /* fixed: nrows elems | STR: Buf | ARR_*: elem values
                            or Buf | LC_STR: keys | NULLABLE: Values           */
  uint64_t data_bytes;
  const dfgpu_pos* pos;  /* STR: nrows entries; ARR_STR/ARR_LC_STR: nelems;
                            NULL when data is the native-encoded stream
                            (dfgpu_cfg.native_strings)                          */
  const uint64_t* arr_offsets; /* ARR_*: nrows cumulative element END offsets  */
  uint64_t nelems;       /* ARR_*: total element count                         */
  const uint8_t* nulls;  /* NULLABLE_*: nrows bytes, 1 = NULL                  */
  const void* dict_data; /* LC_STR: dictionary Buf                             */
  const dfgpu_pos* dict_pos;
  uint64_t dict_n;
  /* dfgpu_cfg.block_lz4: data / pos / arr_offsets / nulls each point at a stream of
     ClickHouse compressed frames (lz4.cu: 16 B checksum slot left zero + method +
     sizes + payload, one frame per 4 KiB of the buffer) instead of the buffer itself;
     the *_bytes fields are the lengths of those streams, data_raw_bytes the length of
     the uncompressed data buffer.  compressed == 0: plain buffers, fields unused.   */
  uint32_t compressed;
  uint32_t reserved0;
  uint64_t data_raw_bytes;
  uint64_t pos_bytes;
  uint64_t arr_offsets_bytes;
  uint64_t nulls_bytes;
} dfgpu_col;

/* ---- input frame table --------------------------------------------------- */
/* One *receiver.RecvBuffer (receiver.go:69-78) after the 19-byte headers
 * (droplet-message.go:182-246) were stripped by the receiver.                 */
typedef struct dfgpu_frame {
  uint64_t offset;  /* of Buffer[Begin] inside the staging buffer              */
  uint32_t len;     /* End-Begin                                               */
  uint16_t vtap_id; /* RecvBuffer.VtapID                                       */
  uint16_t org_id;  /* RecvBuffer.OrgID  (receiver.go:688-704 already applied) */
  uint32_t team_id; /* RecvBuffer.TeamID                                       */
} dfgpu_frame;

/* ---- dictionary snapshot (SmartEncoding source) --------------------------- */
/* Flat mirror of what the controller pushes with AnalyzerSync
 * (message/trident.proto:385-503: PlatformData{interfaces,cidrs,gprocess_infos,
 * peer_connections}, Groups.svcs, SyncResponse{vtap_ips,pod_ips,analyzer_config}),
 * with IP strings already parsed to binary by the shim (net.ParseIP).
 * Build semantics follow server/libs/grpc/grpc_platformdata.go:1090-2028 and
 * grpc_service.go:325-534.                                                     */
typedef struct dfgpu_ip {
  uint8_t is_v6;       /* 0: ip4 valid, 1: ip6 valid                           */
  uint8_t valid;       /* 0: string failed to parse / empty                    */
  uint8_t pad[2];
  uint32_t ip4;        /* host order, a.b.c.d = a<<24|...                      */
  uint8_t ip6[16];
} dfgpu_ip;

typedef struct dfgpu_ip_resource { /* trident.IpResource                      */
  dfgpu_ip ip;
  uint32_t subnet_id;
} dfgpu_ip_resource;

typedef struct dfgpu_interface {   /* trident.Interface                       */
  uint32_t device_type, device_id, if_type, epc_id;
  uint32_t launch_server_id, region_id;
  uint64_t mac;
  uint32_t pod_node_id, az_id, pod_group_id, pod_ns_id, pod_id, pod_cluster_id;
  uint32_t vtap_id, pod_group_type;
  uint32_t ip_res_off, ip_res_cnt; /* slice of ip_resources[]                  */
} dfgpu_interface;

typedef struct dfgpu_cidr {        /* trident.Cidr                            */
  dfgpu_ip prefix;                 /* network address (already masked)         */
  uint32_t prefix_len;
  uint32_t type;                   /* 1 WAN, 2 LAN                             */
  int32_t epc_id;
  uint32_t subnet_id, region_id, az_id;
} dfgpu_cidr;

typedef struct dfgpu_gprocess {    /* trident.GProcessInfo                    */
  uint32_t gprocess_id, vtap_id, pod_id, pid;
} dfgpu_gprocess;

typedef struct dfgpu_vtap_ip {     /* trident.VtapIp                          */
  uint32_t vtap_id, epc_id;
  dfgpu_ip ip;
  uint32_t pod_cluster_id, team_id;
} dfgpu_vtap_ip;

typedef struct dfgpu_pod_ip {      /* trident.PodIp                           */
  uint32_t pod_id, epc_id;
  dfgpu_ip ip;                     /* pod ip; .valid=0 when the string is ""   */
  dfgpu_ip pod_node_ip;
  uint32_t pod_cluster_id, pod_ns_id, pod_group_id, pod_group_type;
} dfgpu_pod_ip;

typedef struct dfgpu_service {     /* trident.ServiceInfo                     */
  uint32_t type;                   /* 1 NODE, 2 POD_GROUP, 3 IP, 4 LB, 7 CUSTOM */
  uint32_t id, pod_cluster_id, epc_id, protocol; /* 0 ANY,1 TCP,2 UDP          */
  /* slices of u32_pool[] / ip_pool[]                                          */
  uint32_t pod_group_ids_off, pod_group_ids_cnt;
  uint32_t ips_off, ips_cnt;
  uint32_t server_ports_off, server_ports_cnt;
  uint32_t pod_service_ids_off, pod_service_ids_cnt;
  uint32_t chost_ids_off, chost_ids_cnt;
  uint32_t host_ids_off, host_ids_cnt;
  uint32_t pod_ids_off, pod_ids_cnt;
} dfgpu_service;

typedef struct dfgpu_dict_snapshot {
  uint64_t version;
  uint32_t region_id;    /* AnalyzerConfig.region_id of this org               */
  uint32_t analyzer_id;  /* AnalyzerConfig.analyzer_id                         */
  const dfgpu_interface* interfaces;      uint64_t n_interfaces;
  const dfgpu_ip_resource* ip_resources;  uint64_t n_ip_resources;
  const dfgpu_cidr* cidrs;                uint64_t n_cidrs;
  const dfgpu_gprocess* gprocesses;       uint64_t n_gprocesses;
  const dfgpu_vtap_ip* vtap_ips;          uint64_t n_vtap_ips;
  const dfgpu_pod_ip* pod_ips;            uint64_t n_pod_ips;
  const dfgpu_service* services;          uint64_t n_services;
  const uint32_t* u32_pool;               uint64_t n_u32_pool;
  const dfgpu_ip* ip_pool;                uint64_t n_ip_pool;
} dfgpu_dict_snapshot;

/* Geo table (server/libs/geo: GEO_ENTRIES + REGION_NAMES of the generated
 * ip_info.go): closed IPv4 ranges -> region code; names[0] is the "unknown"
 * entry (ip_info.py:41).                                                      */
typedef struct dfgpu_geo_entry {
  uint32_t ip_start, ip_end;
  uint8_t region;
  uint8_t pad[3];
} dfgpu_geo_entry;

/* ---- configuration ------------------------------------------------------- */
typedef struct dfgpu_cfg {
  uint64_t max_batch_bytes;    /* largest staging buffer of one submit; 0 => 256 MiB
                                  (1024 frames x 256 KiB, decoder.go:56 x
                                  uniform_sender.rs:159)                        */
  uint32_t max_batch_records;  /* 0 => max_batch_bytes / 64                     */
  uint32_t flow_tag_cache_flush_timeout; /* s; 0 => 1800 (ingester/config/config.go:59) */
  uint32_t flow_tag_cache_max_size;      /* 0 => 1<<18 (config.go:60); the device
                                  set holds 6x this many 32-byte slots (field +
                                  field_value + app_service keys at load <= 1/2)
                                  and reclaims them by expiry / oldest-first
                                  eviction (dfgpu_counters.flow_tag_evicted)      */
  uint32_t trace_id_index_disabled;      /* TraceIdWithIndex.Disabled (config.go:201) */
  uint32_t native_strings;     /* SURVEY 8(f) N1, first step.  1 => the String and Array(String)
                                  columns of l7_flow_log leave the GPU in ClickHouse NATIVE
                                  column encoding (per value: uvarint length, then the bytes =
                                  what ch-go ColStr.EncodeColumn appends) instead of the ch-go
                                  container (Buf + 16-byte Position per value): dfgpu_col.pos is
                                  NULL, data/data_bytes is the encoded stream, arr_offsets stay.
                                  The Position arrays never cross PCIe (-45 % D2H bytes per row).
                                  LC_STR_ROWS / ARR_LC_STR columns carry the same stream; the shim
                                  walks its length prefixes to Append into the LowCardinality column.
                                  The Go side wraps such a column in a proto.ColInput whose
                                  EncodeColumn appends data verbatim.        */
  uint32_t rollup_tables;      /* bit t: keep the raw meters of flow_metrics table DFGPU_TABLE_METRICS_BASE + t
                                  beside its packed columns (device only, 8 B per meter field per row) so
                                  that dfgpu_rollup_add / dfgpu_metrics_reduce can fold them; 0 => none   */
  uint32_t rollup_max_groups;  /* capacity of one table's rollup accumulator in groups; 0 => 1<<20       */
  uint32_t block_lz4;          /* 1: the column buffers of l4_flow_log / l7_flow_log leave the GPU as ClickHouse
                                  compressed frames (LZ4 block per 4 KiB, lz4.cu) — see dfgpu_col.compressed.
                                  Fewer bytes over PCIe and no CPU compression in the writer; the host only
                                  fills each frame's CityHash128 checksum slot.        */
  uint32_t reserved[3];
} dfgpu_cfg;

/* decoder.Counter (decoder.go:60-80) for one submit; the shim adds them up.   */
typedef struct dfgpu_counters {
  int64_t raw_count;   /* RawCount: frames                                     */
  int64_t count;       /* Count                                                */
  int64_t error_count; /* ErrorCount                                           */
  int64_t drop_count;  /* DropCount (always 0: throttle disabled)              */
  int64_t l7_http_count, l7_dns_count, l7_sql_count, l7_nosql_count,
      l7_rpc_count, l7_mq_count;
  int64_t new_field_count, new_field_value_count; /* flow_tag_writer.go:137-146 */
  int64_t rows[DFGPU_TABLE_MAX];
  /* device timing of the kernel pipeline of this submit, in microseconds       */
  double device_us;
  /* flow_tag set pressure (cumulative since dfgpu_init).  The device set holds 6x
   * flow_tag_cache_max_size slots; when more than half full it is rebuilt between
   * batches: expired keys are dropped (exact: an expired key is emitted again like an
   * absent one), then the oldest by lru time (the reference LRU evicts its least
   * recently used entry, libs/lru/lru.go:52-57; evicted keys are emitted again when they
   * return).  A key that does not fit a saturated set is emitted without being
   * recorded (duplicate rows are idempotent in ClickHouse, flow_tag.go:153-160).      */
  int64_t flow_tag_unrecorded;
  int64_t flow_tag_evicted;
} dfgpu_counters;

typedef struct dfgpu_ctx dfgpu_ctx;
typedef struct dfgpu_batch dfgpu_batch;

/* One ctx per decoder goroutine (flow_log.go:226-247: a decoder exclusively owns
 * its platform table view, flow_tag cache and scratch).  Not thread-safe except
 * dfgpu_dict_load / dfgpu_geo_load, which double-buffer and swap between submits. */
int dfgpu_init(int device, const dfgpu_cfg* cfg, dfgpu_ctx** out);
int dfgpu_close(dfgpu_ctx* ctx);
const char* dfgpu_last_error(const dfgpu_ctx* ctx);

/* Replaces PlatformInfoTable.ReloadSlave + ServiceTable swap
 * (grpc_platformdata.go:1166-1228): builds the device tables of one org from a
 * snapshot and swaps them in atomically before the next submit.                */
int dfgpu_dict_load(dfgpu_ctx* ctx, uint16_t org_id, const dfgpu_dict_snapshot* snap);
/* Replaces geo.NewNetmaskGeoTree (flow_log/geo/geo.go:25).                     */
int dfgpu_geo_load(dfgpu_ctx* ctx, const dfgpu_geo_entry* entries, uint64_t n,
                   const char* const* region_names, uint32_t n_names);

/* Replaces the body of Decoder.Run's inner loop for one Gets() batch
 * (decoder.go:161-197, handleTaggedFlow :202, handleProtoLog :220) or
 * Unmarshaller.QueueProcess (unmarshaller.go:226-280).
 *   id_counter       in/out, replaces the process-global L4FlowCounter /
 *                    L7FlowLogCounter (l4_flow_log.go:947, l7_flow_log.go:636)
 *   trace_id_index   in/out, replaces lastTraceIdIndex (l7_flow_log.go:294); may
 *                    be NULL for non-L7 message types
 * frames points to host memory (pageable or pinned).                           */
int dfgpu_submit(dfgpu_ctx* ctx, int msg_type, const uint8_t* frames, size_t nbytes,
                 const dfgpu_frame* tbl, int nframes, uint32_t* id_counter,
                 uint64_t* trace_id_index, dfgpu_batch** out);

/* The same pipeline split in its three phases, for callers that keep inputs
 * resident in HBM (bench "value" leg) or overlap several contexts:
 *   stage : H2D of frames + frame table (async on the ctx stream)
 *   run   : all kernels (async: no allocation, no synchronisation, no D2H — the
 *           record count and everything sized by it stay on the device; every
 *           buffer was sized at dfgpu_init from dfgpu_cfg.max_batch_*)
 *   fetch : D2H of the sizes, then of the column blocks + counters, stream sync.
 *           A batch that outgrew the configured capacities is detected here:
 *           buffers grow and the batch is re-run (never in steady state).
 * frames may be device memory (8-byte aligned); the frame table is host memory. */
int dfgpu_stage(dfgpu_ctx* ctx, int msg_type, const uint8_t* frames, size_t nbytes,
                const dfgpu_frame* tbl, int nframes);
int dfgpu_run(dfgpu_ctx* ctx, uint32_t* id_counter, uint64_t* trace_id_index);
int dfgpu_fetch(dfgpu_ctx* ctx, dfgpu_batch** out);
int dfgpu_sync(dfgpu_ctx* ctx);
/* cudaStream_t of the ctx, for callers that record their own CUDA events.      */
void* dfgpu_stream(dfgpu_ctx* ctx);
/* number of kernels this ctx has launched since init                           */
uint64_t dfgpu_launch_count(const dfgpu_ctx* ctx);
/* Measurement hooks: CUDA events recorded on the ctx stream right around the
 * pack kernels (string / array columns) of every dfgpu_run (<= 256 runs between reads).
 * dfgpu_profile_read synchronises, sums and resets.                               */
int dfgpu_profile_enable(dfgpu_ctx* ctx, int on);
int dfgpu_profile_read(dfgpu_ctx* ctx, double* emit_ms, uint64_t* launches, uint64_t* in_bytes);
/* stage breakdown of the profiled runs, ms[7]: framing, prepare, decode pass, scans, pack (string / array
 * columns), SmartEncoding, flow_tag / app_service tags (all three message types)                          */
int dfgpu_profile_stages(dfgpu_ctx* ctx, double* ms7, uint64_t* runs);
int dfgpu_flush(dfgpu_ctx* ctx);

/* Column block of one table of a batch, in ToInput order.  cols may be NULL to
 * query *ncols / *nrows only.                                                  */
int dfgpu_batch_columns(const dfgpu_batch* b, int table, dfgpu_col* cols, int* ncols,
                        uint64_t* nrows);
int dfgpu_batch_counters(const dfgpu_batch* b, dfgpu_counters* out);
void dfgpu_batch_release(dfgpu_batch* b);

/* Multi-GPU tag merge.  Records shard across GPUs by vtap_id; every decoder
 * keeps its own flow_tag "seen" set (like the reference's per-decoder caches,
 * flow_tag_writer.go:65-77).  After a batch each rank publishes the 128-bit lru-key
 * fingerprints of the flow_tag rows it emitted (row order, 2 x u64 per row; tables
 * L7_CUSTOM_FIELD, L7_CUSTOM_FIELD_VALUE, L7_APP_SERVICE), the ranks all-gather them
 * (NCCL), and every rank imports the others' keys so that later batches do not emit
 * them again.  The lowest rank that emitted a key in a step owns the row; the others
 * drop theirs (host side, deepflow_b200/sharded.py).
 *   fps   : pointer into the batch's host memory, valid until dfgpu_batch_release
 *   times : the `time` column value of the emitting row (lru Add(key, time))          */
int dfgpu_batch_tag_fingerprints(const dfgpu_batch* b, int table, const uint64_t** fps, uint64_t* n);
int dfgpu_tags_import(dfgpu_ctx* ctx, const uint64_t* fps, const uint32_t* times, uint64_t n);
/* Forgets every flow_tag key of this context (a restarted decoder starts with empty lru caches,
 * flow_tag_writer.go:65-77); asynchronous on the ctx stream.                                      */
int dfgpu_tags_reset(dfgpu_ctx* ctx);

/* flow_metrics rollup (BASELINE config 3: 1s -> 1m).  A persistent device
 * accumulator per flow_metrics table, keyed by (time bucket, every tag column of the
 * table) and holding RAW meters, merged with the per-field rules of
 * FlowMeter.ConcurrentMerge (flow-metrics/flow_meter.go:91-99, basic_meter.go:94-113,
 * 277-315,417-424,556-577: wrapping sums, maxes, OooTx/OooRx assigned from the merged
 * row, the derived *Fail fields kept from the first row).
 *   dfgpu_rollup_add      folds the rows of `table` of the last dfgpu_submit (metrics) —
 *                         they stay resident in HBM until the next submit; the table
 *                         must be enabled in dfgpu_cfg.rollup_tables.  A 1s -> 1m
 *                         rollup calls it after every metrics submit of the minute.
 *   dfgpu_rollup_add_rows folds decoded rows the caller already holds in HBM.
 *   dfgpu_rollup_flush    returns every group as table DFGPU_TABLE_METRICS_REDUCED
 *                         (same columns as `table`), in first-appearance order of the
 *                         groups across all adds, and resets the accumulator.
 *   dfgpu_metrics_reduce  = dfgpu_rollup_add + dfgpu_rollup_flush.
 * `table` is one of DFGPU_TABLE_METRICS_BASE+0..8; bucket_seconds = 60 for 1s -> 1m.   */
typedef struct dfgpu_rollup_rows {
  uint64_t nrows;               /* < 2^32                                                */
  const void* const* tag_cols;  /* HOST array of ntag DEVICE pointers: the tag columns of
                                   the table in ToInput order, ch-go container layout
                                   (fixed-width kinds only; observation_point = the
                                   canonical tap-side code, 1 byte)                      */
  int32_t ntag;
  const uint64_t* raw;          /* DEVICE [nrows][nfields]: raw meter fields in struct
                                   order (basic_meter.go:26-44,195-219,386-393,470-496,
                                   691-693; dfgpu_rollup_nfields)                         */
} dfgpu_rollup_rows;
int dfgpu_rollup_add(dfgpu_ctx* ctx, int table, uint32_t bucket_seconds);
int dfgpu_rollup_add_rows(dfgpu_ctx* ctx, int table, uint32_t bucket_seconds, const dfgpu_rollup_rows* rows);
int dfgpu_rollup_flush(dfgpu_ctx* ctx, int table, dfgpu_batch** out);
/* layout of a table's rollup input: number of tag columns / raw meter fields, bytes per value of every tag
 * column and its kind (0 fixed, 1 IPv6, 2 observation_point code, 3/4 strings); arrays of >= 72 entries    */
int dfgpu_rollup_describe(dfgpu_ctx* ctx, int table, int* ntag, int* nfields, uint8_t* elem_sizes, uint8_t* kinds);
int dfgpu_metrics_reduce(dfgpu_ctx* ctx, int table, uint32_t bucket_seconds, dfgpu_batch** out);

#ifdef __cplusplus
}
#endif
#endif /* DFGPU_H */
