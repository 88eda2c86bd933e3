/*
 * gpudecoder — the cgo shim that puts libdfgpu.so in the slot of the flow_log decoder goroutine.
 *
 * Drop this package into server/ingester/gpudecoder of deepflowio/deepflow (it imports only packages that
 * exist there) and construct gpudecoder.NewDecoder where flow_log.go constructs decoder.NewDecoder
 * (server/ingester/flow_log/flow_log.go:226-247); INTEGRATION.md lists the one-line changes around it.
 *
 * Same contract as decoder.Decoder (server/ingester/flow_log/decoder/decoder.go:86-199):
 *   - pulls up to BUFFER_SIZE items per Gets() from its queue.QueueReader; nil = flush tick
 *   - releases every *receiver.RecvBuffer exactly once (receiver.ReleaseRecvBuffer)
 *   - never returns errors upward: counts them (Counter.ErrorCount / DropCount)
 *   - one goroutine owns one dfgpu_ctx (its dictionary view, flow_tag set, counters)
 * Output: one pre-built column block per (table, org) per batch, handed to a ckwriter that accepts blocks
 * (Cache.AddBlock, INTEGRATION.md section 3) instead of rows.
 *
 * This file cannot be compiled in the image it was written in (no Go toolchain): it is written against the
 * reference tree at the commit under /root/reference and kept free of anything that needs code generation.
 */
package gpudecoder

/*
#cgo CFLAGS: -I${SRCDIR}/../../../../include
#cgo LDFLAGS: -ldfgpu
#include <stdlib.h>
#include "dfgpu.h"
#include "dfgpu_wire.h"
*/
import "C"

import (
	"fmt"
	"strconv"
	"sync/atomic"
	"time"
	"unsafe"

	logging "github.com/op/go-logging"

	"github.com/deepflowio/deepflow/server/ingester/common"
	"github.com/deepflowio/deepflow/server/libs/datatype"
	"github.com/deepflowio/deepflow/server/libs/grpc"
	"github.com/deepflowio/deepflow/server/libs/queue"
	"github.com/deepflowio/deepflow/server/libs/receiver"
	"github.com/deepflowio/deepflow/server/libs/stats"
	"github.com/deepflowio/deepflow/server/libs/utils"
)

var log = logging.MustGetLogger("flow_log.gpudecoder")

const BUFFER_SIZE = 1024 // decoder.go:56

// Counter mirrors decoder.Counter (decoder.go:60-80) so that dashboards keep their series.
type Counter struct {
	RawCount     int64 `statsd:"raw-count"`
	L7HTTPCount  int64 `statsd:"l7-http-count"`
	L7DNSCount   int64 `statsd:"l7-dns-count"`
	L7SQLCount   int64 `statsd:"l7-sql-count"`
	L7NoSQLCount int64 `statsd:"l7-nosql-count"`
	L7RPCCount   int64 `statsd:"l7-rpc-count"`
	L7MQCount    int64 `statsd:"l7-mq-count"`
	ErrorCount   int64 `statsd:"err-count"`
	Count        int64 `statsd:"count"`
	DropCount    int64 `statsd:"drop-count"`

	FlowTagUnrecorded int64 `statsd:"flow-tag-unrecorded"`
	FlowTagEvicted    int64 `statsd:"flow-tag-evicted"`

	TotalTime int64 `statsd:"total-time"`
	AvgTime   int64 `statsd:"avg-time"`
}

// Block is one table's column block of one batch: the dfgpu_col array points into pinned memory owned by
// the batch; Release returns it to the library (after the INSERT went out).
type Block struct {
	Table  int
	OrgID  uint16
	Rows   uint64
	Cols   []C.dfgpu_col
	batch  *C.dfgpu_batch
	shared *int32 // blocks of one batch share one release
}

// BlockWriter is what ckwriter.Cache implements after the AddBlock change (INTEGRATION.md section 3).
type BlockWriter interface {
	PutBlock(b *Block)
}

type Decoder struct {
	index        int
	msgType      datatype.MessageType
	platformData *grpc.PlatformInfoTable
	inQueue      queue.QueueReader
	writer       BlockWriter

	ctx          *C.dfgpu_ctx
	staging      []byte // pinned by the library on first use would be better; plain memory works
	frames       []C.dfgpu_frame
	idCounter    C.uint32_t // L4FlowCounter / L7FlowLogCounter (l4_flow_log.go:947, l7_flow_log.go:636)
	traceIdIndex C.uint64_t // lastTraceIdIndex (l7_flow_log.go:294)
	dictVersion  map[uint16]uint64

	counter *Counter
	utils.Closable
}

// Config is the subset of flow_log/config.Config the GPU path reads.
type Config struct {
	Device                   int
	MaxBatchBytes            uint64
	FlowTagCacheFlushTimeout uint32 // ingester/config/config.go:59
	FlowTagCacheMaxSize      uint32 // config.go:60
	TraceIdIndexDisabled     bool   // flow_log/config TraceIdWithIndex.Disabled
	NativeStrings            bool
	BlockLZ4                 bool // column buffers arrive as ClickHouse compressed frames (dfgpu_cfg.block_lz4)
}

func NewDecoder(index int, msgType datatype.MessageType, platformData *grpc.PlatformInfoTable,
	inQueue queue.QueueReader, writer BlockWriter, cfg *Config) (*Decoder, error) {
	var c C.dfgpu_cfg
	c.max_batch_bytes = C.uint64_t(cfg.MaxBatchBytes)
	c.flow_tag_cache_flush_timeout = C.uint32_t(cfg.FlowTagCacheFlushTimeout)
	c.flow_tag_cache_max_size = C.uint32_t(cfg.FlowTagCacheMaxSize)
	if cfg.TraceIdIndexDisabled {
		c.trace_id_index_disabled = 1
	}
	if cfg.NativeStrings {
		c.native_strings = 1
	}
	if cfg.BlockLZ4 {
		c.block_lz4 = 1
	}
	d := &Decoder{index: index, msgType: msgType, platformData: platformData, inQueue: inQueue, writer: writer,
		dictVersion: make(map[uint16]uint64), counter: &Counter{}}
	if rc := C.dfgpu_init(C.int(cfg.Device), &c, &d.ctx); rc != 0 {
		return nil, fmt.Errorf("dfgpu_init(device %d) failed: %d (no CUDA device? there is no CPU fallback)", cfg.Device, int(rc))
	}
	return d, nil
}

func (d *Decoder) GetCounter() interface{} {
	var counter *Counter
	counter, d.counter = d.counter, &Counter{}
	if counter.Count > 0 {
		counter.AvgTime = counter.TotalTime / counter.Count
	}
	return counter
}

// LoadDict pushes one org's platform data (grpc.PlatformInfoTable snapshot flattened by dictsnapshot.go, see
// INTEGRATION.md section 4) when its version changed: the counterpart of PlatformInfoTable.ReloadSlave.
func (d *Decoder) LoadDict(org uint16, version uint64, snap *C.dfgpu_dict_snapshot) error {
	if d.dictVersion[org] == version {
		return nil
	}
	if rc := C.dfgpu_dict_load(d.ctx, C.uint16_t(org), snap); rc != 0 {
		return fmt.Errorf("dfgpu_dict_load(org %d): %s", org, C.GoString(C.dfgpu_last_error(d.ctx)))
	}
	d.dictVersion[org] = version
	return nil
}

func (d *Decoder) Run() {
	common.RegisterCountableForIngester("decoder", d, stats.OptionStatTags{
		"thread": strconv.Itoa(d.index), "msg_type": d.msgType.String(), "impl": "gpu"})
	buffer := make([]interface{}, BUFFER_SIZE)
	byOrg := make(map[uint16][]*receiver.RecvBuffer)
	for !d.Closed() {
		n := d.inQueue.Gets(buffer)
		start := time.Now()
		flush := false
		for i := 0; i < n; i++ {
			if buffer[i] == nil { // flush tick (queue.go:103-111)
				flush = true
				continue
			}
			d.counter.RawCount++
			rb, ok := buffer[i].(*receiver.RecvBuffer)
			if !ok {
				log.Warning("get decode queue data type wrong")
				continue
			}
			byOrg[rb.OrgID] = append(byOrg[rb.OrgID], rb) // one org per submit (DFGPU_EMIXEDORG otherwise)
		}
		for org, bufs := range byOrg {
			d.submit(org, bufs)
			for _, rb := range bufs {
				receiver.ReleaseRecvBuffer(rb) // exactly once, after the bytes were staged (decoder.go:196)
			}
			delete(byOrg, org)
		}
		if flush {
			C.dfgpu_flush(d.ctx)
		}
		d.counter.TotalTime += int64(time.Since(start))
	}
	C.dfgpu_close(d.ctx)
}

// submit concatenates Buffer[Begin:End] of the RecvBuffers into the staging buffer, describes each with a
// dfgpu_frame and makes the one call that replaces handleTaggedFlow / handleProtoLog (decoder.go:202-234).
func (d *Decoder) submit(org uint16, bufs []*receiver.RecvBuffer) {
	total := 0
	for _, rb := range bufs {
		total += rb.End - rb.Begin
	}
	if cap(d.staging) < total {
		d.staging = make([]byte, 0, total+total/4)
	}
	d.staging = d.staging[:0]
	d.frames = d.frames[:0]
	for _, rb := range bufs {
		var f C.dfgpu_frame
		f.offset = C.uint64_t(len(d.staging))
		f.len = C.uint32_t(rb.End - rb.Begin)
		f.vtap_id = C.uint16_t(rb.VtapID)
		f.org_id = C.uint16_t(rb.OrgID) // receiver.parseOrgIdTeamId already applied (receiver.go:688-704)
		f.team_id = C.uint32_t(rb.TeamID)
		d.frames = append(d.frames, f)
		d.staging = append(d.staging, rb.Buffer[rb.Begin:rb.End]...)
	}
	var batch *C.dfgpu_batch
	var framesPtr *C.dfgpu_frame
	var dataPtr *C.uint8_t
	if len(d.frames) > 0 {
		framesPtr = &d.frames[0]
	}
	if len(d.staging) > 0 {
		dataPtr = (*C.uint8_t)(unsafe.Pointer(&d.staging[0]))
	}
	rc := C.dfgpu_submit(d.ctx, C.int(d.msgType), dataPtr, C.size_t(len(d.staging)), framesPtr, C.int(len(d.frames)),
		&d.idCounter, &d.traceIdIndex, &batch)
	if rc != 0 {
		// like a panic-free Unmarshal error on every frame: counted, never returned upward
		d.counter.ErrorCount += int64(len(bufs))
		log.Warningf("dfgpu_submit: %s", C.GoString(C.dfgpu_last_error(d.ctx)))
		return
	}
	var cn C.dfgpu_counters
	C.dfgpu_batch_counters(batch, &cn)
	d.counter.Count += int64(cn.count)
	d.counter.ErrorCount += int64(cn.error_count)
	d.counter.DropCount += int64(cn.drop_count)
	d.counter.L7HTTPCount += int64(cn.l7_http_count)
	d.counter.L7DNSCount += int64(cn.l7_dns_count)
	d.counter.L7SQLCount += int64(cn.l7_sql_count)
	d.counter.L7NoSQLCount += int64(cn.l7_nosql_count)
	d.counter.L7RPCCount += int64(cn.l7_rpc_count)
	d.counter.L7MQCount += int64(cn.l7_mq_count)
	d.counter.FlowTagUnrecorded = int64(cn.flow_tag_unrecorded)
	d.counter.FlowTagEvicted = int64(cn.flow_tag_evicted)

	tables := tablesOf(d.msgType)
	shared := new(int32)
	*shared = int32(len(tables))
	for _, t := range tables {
		var ncols C.int
		var nrows C.uint64_t
		if C.dfgpu_batch_columns(batch, C.int(t), nil, &ncols, &nrows) != 0 || nrows == 0 {
			releaseShared(batch, shared)
			continue
		}
		cols := make([]C.dfgpu_col, int(ncols))
		C.dfgpu_batch_columns(batch, C.int(t), &cols[0], &ncols, &nrows)
		d.writer.PutBlock(&Block{Table: t, OrgID: org, Rows: uint64(nrows), Cols: cols, batch: batch, shared: shared})
	}
}

func tablesOf(m datatype.MessageType) []int {
	switch m {
	case datatype.MESSAGE_TYPE_TAGGEDFLOW:
		return []int{C.DFGPU_TABLE_L4_FLOW_LOG}
	case datatype.MESSAGE_TYPE_PROTOCOLLOG:
		return []int{C.DFGPU_TABLE_L7_FLOW_LOG, C.DFGPU_TABLE_L7_CUSTOM_FIELD, C.DFGPU_TABLE_L7_CUSTOM_FIELD_VALUE,
			C.DFGPU_TABLE_L7_APP_SERVICE}
	case datatype.MESSAGE_TYPE_METRICS:
		t := make([]int, 0, 10)
		for i := 0; i < 9; i++ {
			t = append(t, C.DFGPU_TABLE_METRICS_BASE+i)
		}
		return append(t, C.DFGPU_TABLE_METRICS_APP_SERVICE)
	}
	return nil
}

// Release is called by the writer when the block's INSERT is done; the last block of a batch frees it.
func (b *Block) Release() { releaseShared(b.batch, b.shared) }

func releaseShared(batch *C.dfgpu_batch, shared *int32) {
	// the tables of one batch go to different writer goroutines: the last release frees the batch
	// (dfgpu_batch_release is safe from any thread: the host pool is locked)
	if atomic.AddInt32(shared, -1) == 0 {
		C.dfgpu_batch_release(batch)
	}
}

// Frames walks a stream of compressed frames as dfgpu_cfg.block_lz4 returns them for one column buffer
// (dfgpu_col.data / pos / arr_offsets / nulls with compressed == 1): 16 bytes of checksum slot, 1 byte
// method (0x82 LZ4, 0x02 stored), uint32 compressed size (9-byte header included), uint32 uncompressed
// size, payload.  The GPU leaves the checksum slot zero; seal fills it (CityHash128 of header + payload,
// github.com/go-faster/city — the hash ch-go's compress.Writer uses) right before the frame goes to the
// connection, so that the bytes written are exactly what a ClickHouse server expects.
func Frames(stream []byte, seal func(frame []byte), emit func(frame []byte)) error {
	for len(stream) > 0 {
		if len(stream) < 25 {
			return fmt.Errorf("truncated frame header (%d bytes left)", len(stream))
		}
		size := int(uint32(stream[17]) | uint32(stream[18])<<8 | uint32(stream[19])<<16 | uint32(stream[20])<<24)
		if size < 9 || 16+size > len(stream) {
			return fmt.Errorf("frame of %d bytes in a stream of %d", size, len(stream))
		}
		frame := stream[:16+size]
		seal(frame)
		emit(frame)
		stream = stream[16+size:]
	}
	return nil
}
