/*
 * dfgpu_wire.h — SURVEY A1: the droplet message headers in front of every frame the agents send, and the
 * org / team normalisation the receiver applies before a frame reaches a decoder queue.  Host-side, plain C
 * (header only): what the shim needs to turn a received message into a dfgpu_frame.
 *   BaseHeader.Decode     server/libs/datatype/droplet-message.go:182-201  (consts :150-170, types :36-143)
 *   FlowHeader.Decode     server/libs/datatype/droplet-message.go:231-246  (consts :210-219)
 *   parseOrgIdTeamId      server/libs/receiver/receiver.go:688-704         (ckdb/table.go:26-32)
 */
#ifndef DFGPU_WIRE_H
#define DFGPU_WIRE_H

#include <stddef.h>
#include <stdint.h>

#include "dfgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

#define DFGPU_MESSAGE_HEADER_LEN 5   /* u32 BE frame size (headers included) + u8 message type */
#define DFGPU_FLOW_HEADER_LEN 14     /* version, encoder, team, org, reserved, agent id, reserved */
#define DFGPU_MESSAGE_TYPE_MAX 22    /* MESSAGE_TYPE_MAX */
#define DFGPU_FLOW_VERSION_LATEST 0x8000
#define DFGPU_ENCODER_RAW 0
#define DFGPU_ENCODER_ZLIB 1
#define DFGPU_ENCODER_GZIP 2
#define DFGPU_ENCODER_ZSTD 3
#define DFGPU_DEFAULT_ORG_ID 1
#define DFGPU_INVALID_ORG_ID 0
#define DFGPU_MAX_ORG_ID 1024
#define DFGPU_DEFAULT_TEAM_ID 1
#define DFGPU_INVALID_TEAM_ID 0

/* MessageType.HeaderType(): 0 = 4B+1B, 1 = 4B+1B length unchecked, 2 = + flow header, 3 = unknown */
static inline int dfgpu_message_header_type(uint8_t type) {
  if (type >= DFGPU_MESSAGE_TYPE_MAX) return 3;
  if (type == 0) return 0; /* MESSAGE_TYPE_COMPRESS */
  if (type == 1) return 1; /* MESSAGE_TYPE_SYSLOG   */
  return 2;
}

typedef struct dfgpu_base_header {
  uint32_t frame_size; /* whole message, base header included */
  uint8_t type;
} dfgpu_base_header;

/* 0, or DFGPU_EINVAL where BaseHeader.Decode returns an error */
static inline int dfgpu_base_header_decode(const uint8_t* buf, dfgpu_base_header* h) {
  h->frame_size = (uint32_t)buf[0] << 24 | (uint32_t)buf[1] << 16 | (uint32_t)buf[2] << 8 | (uint32_t)buf[3];
  h->type = buf[4];
  switch (dfgpu_message_header_type(h->type)) {
    case 0: return h->frame_size <= DFGPU_MESSAGE_HEADER_LEN ? DFGPU_EINVAL : 0;
    case 1: return 0;
    case 2: return h->frame_size < DFGPU_MESSAGE_HEADER_LEN + DFGPU_FLOW_HEADER_LEN ? DFGPU_EINVAL : 0;
    default: return DFGPU_EINVAL;
  }
}

typedef struct dfgpu_flow_header {
  uint16_t version;
  uint8_t encoder;
  uint32_t team_id;
  uint16_t org_id;
  uint16_t agent_id;
} dfgpu_flow_header;

static inline uint16_t dfgpu_le16(const uint8_t* p) { return (uint16_t)(p[0] | p[1] << 8); }
static inline uint32_t dfgpu_le32(const uint8_t* p) {
  return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
}

/* buf: the 14 bytes behind the base header */
static inline void dfgpu_flow_header_decode(const uint8_t* buf, dfgpu_flow_header* h) {
  h->version = dfgpu_le16(buf + 0);
  if (h->version == DFGPU_FLOW_VERSION_LATEST) {
    h->encoder = buf[2];
    h->team_id = dfgpu_le32(buf + 3);
    h->org_id = dfgpu_le16(buf + 7);
    h->agent_id = dfgpu_le16(buf + 11);
  } else { /* header of agents <= v6.5.8: u32 version, u32 team, u32 org, u16 vtap at offset 12 */
    h->encoder = DFGPU_ENCODER_RAW;
    h->team_id = DFGPU_DEFAULT_TEAM_ID;
    h->org_id = DFGPU_DEFAULT_ORG_ID;
    h->agent_id = dfgpu_le16(buf + 12);
  }
}

/* parseOrgIdTeamId; *invalid += 1 when the org id is out of range (receiver counter.Invalid) */
static inline void dfgpu_org_team_normalize(uint16_t org, uint32_t team, uint16_t* org_out, uint32_t* team_out,
                                            uint64_t* invalid) {
  if (team == DFGPU_INVALID_TEAM_ID) team = DFGPU_DEFAULT_TEAM_ID;
  if (org == DFGPU_INVALID_ORG_ID) {
    org = DFGPU_DEFAULT_ORG_ID;
  } else if (org > DFGPU_MAX_ORG_ID) {
    if (invalid) (*invalid)++;
    org = DFGPU_DEFAULT_ORG_ID;
  }
  *org_out = org;
  *team_out = team;
}

/* One received message (base header first) -> the fields of its dfgpu_frame / RecvBuffer
 * (receiver.go:69-78: Begin = after the headers, End = FrameSize).  payload_off / payload_len locate
 * Buffer[Begin:End] inside msg.  Returns 0, DFGPU_EINVAL (bad header / short buffer) or DFGPU_ENOENT when the
 * message is not one of the three message types of this path.  *encoder tells the caller whether the payload is
 * compressed (DFGPU_ENCODER_*): such frames are inflated before they are staged.                              */
static inline int dfgpu_frame_from_message(const uint8_t* msg, size_t len, dfgpu_frame* frame, int* msg_type,
                                           size_t* payload_off, uint8_t* encoder, uint64_t* invalid) {
  dfgpu_base_header b;
  dfgpu_flow_header f;
  if (len < DFGPU_MESSAGE_HEADER_LEN) return DFGPU_EINVAL;
  if (dfgpu_base_header_decode(msg, &b)) return DFGPU_EINVAL;
  if (b.type != DFGPU_MSG_METRICS && b.type != DFGPU_MSG_TAGGEDFLOW && b.type != DFGPU_MSG_PROTOCOLLOG) return DFGPU_ENOENT;
  if (b.frame_size > len) return DFGPU_EINVAL;
  dfgpu_flow_header_decode(msg + DFGPU_MESSAGE_HEADER_LEN, &f);
  frame->offset = 0;
  frame->len = b.frame_size - (DFGPU_MESSAGE_HEADER_LEN + DFGPU_FLOW_HEADER_LEN);
  frame->vtap_id = f.agent_id;
  dfgpu_org_team_normalize(f.org_id, f.team_id, &frame->org_id, &frame->team_id, invalid);
  if (msg_type) *msg_type = b.type;
  if (payload_off) *payload_off = DFGPU_MESSAGE_HEADER_LEN + DFGPU_FLOW_HEADER_LEN;
  if (encoder) *encoder = f.encoder;
  return 0;
}

#ifdef __cplusplus
}
#endif
#endif /* DFGPU_WIRE_H */
