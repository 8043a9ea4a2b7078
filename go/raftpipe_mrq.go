// raftpipe_mrq.go — drop-in replacement for raft.go + raftpipe.go of chzchzchz/raftsql.
//
// SOURCE ONLY (no Go toolchain in the build image; never compiled).  It keeps the seam of the reference
// byte for byte — `type raftPipe struct{ProposeC, CommitC, ErrorC}`, `NewRaftPipe(id, peers, proposeC)`,
// `(*raftPipe).Close()` (reference raftpipe.go:3-17) — so db.go, httpapi.go and server/main.go compile
// unchanged, and replaces what sits below it: instead of one etcd-raft `raft.Node` per process
// (reference raft.go:62-78,152-165) the node's consensus arithmetic runs in the mrq engine on a B200
// through the cgo binding in ./mrq.  The surrounding duties of reference raftNode are kept on the Go side,
// exactly as the Python host node of this repo does (raftsql_b200/hostnode.py, which IS tested):
//
//   reference                                   here
//   ---------                                   ----
//   rc.node.Tick()             raft.go:224      eng.Tick(slot) after eng.Step(slot, msgs)
//   rc.node.Step(ctx, m)       raft.go:269      inbound messages queued by Process(), posted each tick
//   rc.node.Propose(ctx, b)    raft.go:214      eng.Propose(slot, {0}, {n}); payloads kept in rc.pending
//   <-rc.node.Ready()          raft.go:227      eng.Ready(committed, term, role, out) + out-word decoding
//   rc.wal.Save / Append       raft.go:228-229  unchanged (etcd wal + MemoryStorage hold the entries)
//   rc.transport.Send          raft.go:230      unchanged (rafthttp); messages rebuilt from the out word
//   rc.publishEntries          raft.go:82-96    publish (applied, committed] — on COMMIT, not on append
//   rc.node.Advance()          raft.go:235      nothing to do: the engine has no Ready queue
//
// For many groups per process (the multi-raft shape the engine is built for) use NewMultiRaftPipe
// (multiraftpipe_mrq.go): one engine with G groups, one tick loop, and committed[g] advances demultiplexed
// into per-group CommitC's.
package raftsql

import (
	"time"

	"github.com/chzchzchz/raftsql/go/mrq"
)

// raftPipe: unchanged from reference raftpipe.go:3-7.
type raftPipe struct {
	ProposeC chan<- string
	CommitC  <-chan *string
	ErrorC   <-chan error
}

// NewRaftPipe: same signature as reference raftpipe.go:9-12.
func NewRaftPipe(id int, peers []string, proposeC chan string) *raftPipe {
	cC, eC := newRaftNode(id, peers, proposeC)
	return &raftPipe{ProposeC: proposeC, CommitC: cC, ErrorC: eC}
}

// Close: unchanged from reference raftpipe.go:14-17.
func (rp *raftPipe) Close() error {
	close(rp.ProposeC)
	return <-rp.ErrorC
}

type entry struct {
	term uint64
	data []byte
}

// inbound is a raftpb.Message reduced to what the host and the engine need.
type inbound struct {
	typ, from                   uint8
	term, logterm, index, commit uint64
	reject                      bool
	rejectHint                  uint64
	entries                     []entry
}

type raftNode struct {
	proposeC <-chan string
	commitC  chan *string
	errorC   chan error
	id       int
	peers    []string

	eng     *mrq.Engine
	log     []entry // log[i-1] is entry i (raft.MemoryStorage in the reference, raft.go:70)
	pending [][]byte
	next    []uint64 // Progress.Next per peer id (message construction, host side)
	applied uint64
	term, vote, commit uint64
	role, lead         uint8

	inbox chan inbound // filled by Process() (rafthttp.Raft, reference raft.go:268-270)
	stopc chan struct{}
}

// newRaftNode keeps the contract documented at reference raft.go:57-61: replayed entries, then nil, then
// live entries on commitC; close proposeC and read errorC to shut down.
func newRaftNode(id int, peers []string, proposeC <-chan string) (<-chan *string, <-chan error) {
	rc := &raftNode{
		proposeC: proposeC,
		commitC:  make(chan *string), // unbuffered, like raft.go:65
		errorC:   make(chan error),   // raft.go:66
		id:       id,
		peers:    peers,
		next:     make([]uint64, len(peers)+1),
		inbox:    make(chan inbound, 4096),
		stopc:    make(chan struct{}),
	}
	go rc.startRaft()
	return rc.commitC, rc.errorC
}

func (rc *raftNode) startRaft() {
	eng, err := mrq.New(mrq.Config{Groups: 1, Replicas: uint32(len(rc.peers)), SelfID: uint32(rc.id),
		ElectionTick: 10, HeartbeatTick: 1, InboxSlots: 1}) // raft.go:152-159
	if err != nil {
		close(rc.commitC)
		rc.errorC <- err
		close(rc.errorC)
		return
	}
	rc.eng = eng
	// replayWAL (raft.go:122-134) goes here unchanged: wal.ReadAll -> rc.log, eng.ImportState(HardState,
	// lastIndex, lastTerm), publish the committed prefix, then the nil sentinel:
	rc.commitC <- nil
	// transport.Start()/AddPeer (raft.go:170-184) unchanged; serveRaft (raft.go:248-266) unchanged.
	go rc.serveChannels()
}

// Process implements rafthttp.Raft (reference raft.go:268-270): inbound messages wait for the next tick.
func (rc *raftNode) Process(m inbound) error {
	select {
	case rc.inbox <- m:
	default: // a full mailbox drops the message, as a congested transport would
	}
	return nil
}

func (rc *raftNode) serveChannels() {
	ticker := time.NewTicker(100 * time.Millisecond) // raft.go:207
	defer ticker.Stop()
	props := make(chan string, 1024)
	go func() { // raft.go:211-218
		for p := range rc.proposeC {
			props <- p
		}
		close(rc.stopc)
	}()
	committed, term := make([]uint64, 1), make([]uint64, 1)
	role, out := make([]uint8, 1), make([]uint32, 1)
	for {
		select {
		case <-ticker.C:
			// 1. inbound messages -> engine inbox (MsgApp is resolved against rc.log first: maybeAppend)
			var msgs []mrq.Msg
		drain:
			for {
				select {
				case m := <-rc.inbox:
					msgs = append(msgs, rc.resolve(m))
				default:
					break drain
				}
			}
			// 2. proposals: node.Propose blocks until there is a leader; forward when we are a follower
		collect:
			for {
				select {
				case p := <-props:
					rc.pending = append(rc.pending, []byte(p))
				default:
					break collect
				}
			}
			if err := rc.eng.Step(0, msgs); err != nil {
				rc.writeError(err)
				return
			}
			nprop := 0
			if rc.role == mrq.RoleLeader && len(rc.pending) > 0 {
				nprop = len(rc.pending)
				rc.eng.Propose(0, []uint64{0}, []uint32{uint32(nprop)})
			}
			// 3. the tick: Step + Propose + Tick for the group, one kernel launch
			if err := rc.eng.Tick(0); err != nil {
				rc.writeError(err)
				return
			}
			// 4. Ready: HardState + out word -> wal.Save, transport.Send, publish (raft.go:227-235)
			if err := rc.eng.Ready(committed, term, role, out); err != nil {
				rc.writeError(err)
				return
			}
			rc.handleReady(committed[0], term[0], role[0], out[0], nprop)
		case <-rc.stopc:
			rc.stop()
			return
		}
	}
}

// resolve turns one inbound message into the engine's form; for MsgApp it runs the log-matching half of
// handleAppendEntries here (the log lives on the host) and reports the outcome (include/mrq.h MSG_APP).
func (rc *raftNode) resolve(m inbound) mrq.Msg {
	out := mrq.Msg{Group: 0, From: m.from, Type: m.typ, Term: m.term, Index: m.index, LogTerm: m.logterm, Commit: m.commit}
	if m.reject {
		out.Type |= mrq.MsgReject
	}
	if m.typ != mrq.MsgApp {
		return out
	}
	// maybeAppend: matchTerm(m.index, m.logterm) -> truncate conflicts, append, lastnewi = index+len(entries)
	if m.index > uint64(len(rc.log)) || (m.index > 0 && rc.log[m.index-1].term != m.logterm) {
		out.Type |= mrq.MsgReject
		return out
	}
	for k, e := range m.entries {
		i := m.index + 1 + uint64(k)
		if i <= uint64(len(rc.log)) {
			if rc.log[i-1].term != e.term {
				rc.log = append(rc.log[:i-1], e)
			}
		} else {
			rc.log = append(rc.log, e)
		}
	}
	lastnewi := m.index + uint64(len(m.entries))
	out.Index = uint64(len(rc.log))
	out.LogTerm = rc.log[len(rc.log)-1].term
	if m.commit < lastnewi {
		out.Commit = m.commit
	} else {
		out.Commit = lastnewi
	}
	return out
}

// handleReady mirrors raftsql_b200/hostnode.py HostNode._ready (the tested implementation): append the
// entries the engine accepted (the empty entry of a new term first), persist, rebuild Ready.Messages from the
// out word (MsgVote on OutCampaign, MsgVoteResp from the reply bits, MsgApp/MsgHeartbeat on the broadcast
// flags, MsgAppResp/MsgHeartbeatResp from the ack bits), then publish (applied, committed].
func (rc *raftNode) handleReady(committed, term uint64, role uint8, out uint32, nprop int) {
	if role == mrq.RoleLeader {
		if out&mrq.OutBecameLeader != 0 {
			rc.log = append(rc.log, entry{term: term})
		}
		for _, p := range rc.pending[:nprop] {
			rc.log = append(rc.log, entry{term: term, data: p})
		}
		rc.pending = rc.pending[nprop:]
	}
	rc.term, rc.role, rc.commit = term, role, committed
	// wal.Save(HardState{term, vote, commit}, new entries); transport.Send(messages) — see hostnode.py
	for rc.applied < committed && rc.applied < uint64(len(rc.log)) {
		rc.applied++
		if d := rc.log[rc.applied-1].data; len(d) > 0 { // raft.go:84-86
			s := string(d)
			select {
			case rc.commitC <- &s:
			case <-rc.stopc:
				return
			}
		}
	}
}

func (rc *raftNode) writeError(err error) { // raft.go:136-142
	close(rc.commitC)
	rc.errorC <- err
	close(rc.errorC)
	rc.eng.Close()
}

func (rc *raftNode) stop() { // raft.go:191-196
	close(rc.commitC)
	close(rc.errorC)
	rc.eng.Close()
}
