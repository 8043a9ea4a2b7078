// Package mrq is the cgo binding of libmrq.so (include/mrq.h) — the Blackwell-native multi-raft quorum
// engine.  SOURCE ONLY: there is no Go toolchain in the build image, so this file has never been compiled;
// it is the binding a maintainer of chzchzchz/raftsql would add next to raft.go (see INTEGRATION.md).
//
// cgo rules observed: C never retains a Go pointer after a call returns (every mrq_* entry point copies
// its inputs before returning); outputs are written into Go-owned slices passed for the duration of the
// call; there are no callbacks from C into Go.
package mrq

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../raftsql_b200 -lmrq
#include <stdlib.h>
#include "mrq.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"unsafe"
)

// Message types (raftpb.MessageType numbering of the etcd v2.2/v2.3 era).
const (
	MsgApp           = C.MRQ_MSG_APP
	MsgAppResp       = C.MRQ_MSG_APP_RESP
	MsgVote          = C.MRQ_MSG_VOTE
	MsgVoteResp      = C.MRQ_MSG_VOTE_RESP
	MsgHeartbeat     = C.MRQ_MSG_HEARTBEAT
	MsgHeartbeatResp = C.MRQ_MSG_HEARTBEAT_RESP
	MsgReject        = C.MRQ_MSG_REJECT
)

// Out-word flags (the engine's stand-in for Ready.Messages, reference raft.go:227-230).
const (
	OutCampaign       = C.MRQ_OUT_CAMPAIGN
	OutBecameLeader   = C.MRQ_OUT_BECAME_LEADER
	OutBcastAppend    = C.MRQ_OUT_BCAST_APPEND
	OutBcastHeartbeat = C.MRQ_OUT_BCAST_HEARTBEAT
	OutSteppedDown    = C.MRQ_OUT_STEPPED_DOWN
	OutPropDropped    = C.MRQ_OUT_PROP_DROPPED
	OutPropForward    = C.MRQ_OUT_PROP_FORWARD
	OutCommitAdvanced = C.MRQ_OUT_COMMIT_ADVANCED
	OutVoteReplyShift = C.MRQ_OUT_VOTE_REPLY_SHIFT
	OutAckReplyShift  = C.MRQ_OUT_ACK_REPLY_SHIFT
)

const (
	RoleFollower  = C.MRQ_ROLE_FOLLOWER
	RoleCandidate = C.MRQ_ROLE_CANDIDATE
	RoleLeader    = C.MRQ_ROLE_LEADER
)

// Config mirrors raft.Config as the reference fills it (raft.go:152-159), for G groups.
type Config struct {
	Groups        uint64 // G
	GroupBase     uint64
	Replicas      uint32 // len(peers), raft.go:148
	SelfID        uint32 // raft.Config.ID, raft.go:153
	ElectionTick  uint32 // raft.go:154 (10)
	HeartbeatTick uint32 // raft.go:155 (1)
	Seed          uint64
	Device        int32
	InboxSlots    uint32
}

// Msg is one inbound raftpb.Message reduced to the fields the path reads.
type Msg struct {
	Group, Term, Index, LogTerm, Commit uint64
	Type, From                          uint8
}

// Engine is this node's replica of Config.Groups raft groups, resident on one B200.
// One goroutine drives one Engine, like the single select loop of reference raft.go:221-245.
type Engine struct {
	h *C.mrq_engine
	g uint64
	r uint32
}

func lastErr(h *C.mrq_engine) error { return errors.New(C.GoString(C.mrq_last_error(h))) }

// New is raft.StartNode for G groups (reference raft.go:161-165).
func New(c Config) (*Engine, error) {
	var cfg C.mrq_config
	C.mrq_config_default(&cfg)
	cfg.n_groups = C.uint64_t(c.Groups)
	cfg.group_base = C.uint64_t(c.GroupBase)
	cfg.n_replicas = C.uint32_t(c.Replicas)
	cfg.self_id = C.uint32_t(c.SelfID)
	if c.ElectionTick != 0 {
		cfg.election_tick = C.uint32_t(c.ElectionTick)
	}
	if c.HeartbeatTick != 0 {
		cfg.heartbeat_tick = C.uint32_t(c.HeartbeatTick)
	}
	cfg.seed = C.uint64_t(c.Seed)
	cfg.device = C.int32_t(c.Device)
	if c.InboxSlots != 0 {
		cfg.inbox_slots = C.uint32_t(c.InboxSlots)
	}
	var h *C.mrq_engine
	if rc := C.mrq_create(&cfg, &h); rc != 0 {
		return nil, fmt.Errorf("mrq_create: %d: %s", int(rc), C.GoString(C.mrq_last_error(nil)))
	}
	return &Engine{h: h, g: c.Groups, r: c.Replicas}, nil
}

// Close is node.Stop() (reference raft.go:141,195).
func (e *Engine) Close() {
	if e.h != nil {
		C.mrq_destroy(e.h)
		e.h = nil
	}
}

// Step posts the tick's inbound messages (node.Step, reference raft.go:268-270), sparse form.
func (e *Engine) Step(slot uint32, msgs []Msg) error {
	if len(msgs) == 0 {
		if rc := C.mrq_clear_inbox(e.h, C.uint32_t(slot)); rc != 0 {
			return lastErr(e.h)
		}
		return nil
	}
	buf := make([]C.mrq_msg, len(msgs))
	for i, m := range msgs {
		buf[i].group, buf[i].term, buf[i].index = C.uint64_t(m.Group), C.uint64_t(m.Term), C.uint64_t(m.Index)
		buf[i].logterm, buf[i].commit = C.uint64_t(m.LogTerm), C.uint64_t(m.Commit)
		buf[i]._type, buf[i].from = C.uint8_t(m.Type), C.uint8_t(m.From)
	}
	if rc := C.mrq_post_inbox_delta(e.h, C.uint32_t(slot), &buf[0], C.size_t(len(buf)), 0); rc != 0 {
		return lastErr(e.h)
	}
	return nil
}

// Propose is node.Propose (reference raft.go:211-215): count entries per group; payloads stay in Go.
func (e *Engine) Propose(slot uint32, groups []uint64, counts []uint32) error {
	if len(groups) == 0 {
		return nil
	}
	rc := C.mrq_propose(e.h, C.uint32_t(slot), (*C.uint64_t)(unsafe.Pointer(&groups[0])),
		(*C.uint32_t)(unsafe.Pointer(&counts[0])), C.size_t(len(groups)))
	if rc != 0 {
		return lastErr(e.h)
	}
	return nil
}

// Tick is node.Tick() for every group (reference raft.go:223-224) fused with the Step of the posted
// inbox: one sm_100a kernel launch.
func (e *Engine) Tick(slot uint32) error {
	if rc := C.mrq_tick(e.h, C.uint32_t(slot)); rc != 0 {
		return lastErr(e.h)
	}
	return nil
}

// Ready drains what <-node.Ready() would carry (reference raft.go:227): HardState{Term,Commit} and the
// per-group out word.  Slices must have Groups elements.
func (e *Engine) Ready(committed, term []uint64, role []uint8, out []uint32) error {
	var pc, pt *C.uint64_t
	var pr *C.uint8_t
	if committed != nil {
		pc = (*C.uint64_t)(unsafe.Pointer(&committed[0]))
	}
	if term != nil {
		pt = (*C.uint64_t)(unsafe.Pointer(&term[0]))
	}
	if role != nil {
		pr = (*C.uint8_t)(unsafe.Pointer(&role[0]))
	}
	if rc := C.mrq_sync_commits(e.h, pc, pr, pt); rc != 0 {
		return lastErr(e.h)
	}
	if out != nil {
		if rc := C.mrq_sync_out(e.h, (*C.uint32_t)(unsafe.Pointer(&out[0]))); rc != 0 {
			return lastErr(e.h)
		}
	}
	return nil
}

// QuorumCommit runs the standalone quorum kernel (maybeCommit on every leader group).
func (e *Engine) QuorumCommit() error {
	if rc := C.mrq_quorum_commit(e.h); rc != 0 {
		return lastErr(e.h)
	}
	return nil
}
