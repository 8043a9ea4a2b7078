// multiraftpipe_mrq.go — the multi-group variant of the seam (SURVEY §8b, §8f row f1).
//
// SOURCE ONLY (no Go toolchain in the build image; never compiled).  The tested implementation of exactly
// this design is raftsql_b200/multipipe.py (tests/test_multipipe_cpu.py, tests/test_zz_multipipe_gpu.py).
//
// One process hosts this node's replica of G raft groups over ONE engine: the per-tick arithmetic of all
// groups is a single eng.Tick (one kernel pass), and the committed[g] advances the engine reports are
// demultiplexed into per-group CommitC's.  Each group keeps the raftPipe protocol of reference
// raftpipe.go:3-17 / raft.go:57-61: replayed entries, then nil, then live entries in log order; closing
// every ProposeC shuts the node down; ErrorC carries at most one error and is then closed.
package raftsql

import (
	"time"

	"github.com/chzchzchz/raftsql/go/mrq"
)

// MultiRaftPipe: G raftPipes that share one node loop and one error channel.
type MultiRaftPipe struct {
	ProposeC []chan<- string
	CommitC  []<-chan *string
	ErrorC   <-chan error
}

// groupState is what raftNode keeps per group in the single-group shim (raftpipe_mrq.go): the log, the
// proposals waiting for a leader, Progress.Next, and the applied index.
type groupState struct {
	log     []entry
	pending [][]byte
	posted  int // proposals handed to the engine this tick
	next    []uint64
	applied uint64
	term    uint64
	role    uint8
}

type multiRaftNode struct {
	id       int
	peers    []string
	eng      *mrq.Engine
	groups   []groupState
	proposeC []chan string
	commitC  []chan *string
	errorC   chan error
	inbox    chan groupInbound // filled by Process(): rafthttp delivers (group, message)
	stopc    chan struct{}
}

type groupInbound struct {
	group uint64
	m     inbound
}

// NewMultiRaftPipe starts this node's replicas of nGroups groups.  proposeC[g] is the caller's proposal
// channel for group g (server/main.go:30 makes one; a multi-raft host makes G).
func NewMultiRaftPipe(id int, peers []string, proposeC []chan string) *MultiRaftPipe {
	n := len(proposeC)
	rc := &multiRaftNode{
		id: id, peers: peers, groups: make([]groupState, n), proposeC: proposeC,
		commitC: make([]chan *string, n), errorC: make(chan error), // unbuffered, like raft.go:65-66
		inbox: make(chan groupInbound, 1<<16), stopc: make(chan struct{}),
	}
	mp := &MultiRaftPipe{ErrorC: rc.errorC}
	for g := range proposeC {
		rc.commitC[g] = make(chan *string)
		rc.groups[g].next = make([]uint64, len(peers)+1)
		mp.ProposeC = append(mp.ProposeC, proposeC[g])
		mp.CommitC = append(mp.CommitC, rc.commitC[g])
	}
	go rc.run()
	return mp
}

// Close: raftpipe.go:14-17 for every group.
func (mp *MultiRaftPipe) Close() error {
	for _, c := range mp.ProposeC {
		close(c)
	}
	return <-mp.ErrorC
}

// Process implements rafthttp.Raft for (group, message) pairs (reference raft.go:268-270).
func (rc *multiRaftNode) Process(group uint64, m inbound) error {
	select {
	case rc.inbox <- groupInbound{group, m}:
	default: // a full mailbox drops the message, as a congested transport would
	}
	return nil
}

func (rc *multiRaftNode) run() {
	G := len(rc.groups)
	eng, err := mrq.New(mrq.Config{Groups: uint64(G), Replicas: uint32(len(rc.peers)), SelfID: uint32(rc.id),
		ElectionTick: 10, HeartbeatTick: 1, InboxSlots: 1}) // raft.go:152-159, for G groups at once
	if err != nil {
		rc.fail(err)
		return
	}
	rc.eng = eng
	// replayWAL per group (raft.go:122-134) -> one eng.ImportState of the restored columns, publish each
	// group's committed prefix, then the nil sentinel on every CommitC:
	for g := range rc.commitC {
		rc.commitC[g] <- nil
	}
	// proposals: one goroutine per group is the Go way (raft.go:211-218); the node stops when all are closed
	type prop struct {
		g int
		s string
	}
	props := make(chan prop, 4096)
	open := make(chan int, G)
	for g, pc := range rc.proposeC {
		go func(g int, pc <-chan string) {
			for p := range pc {
				props <- prop{g, p}
			}
			open <- g
		}(g, pc)
	}
	go func() {
		for k := 0; k < G; k++ {
			<-open
		}
		close(rc.stopc)
	}()

	ticker := time.NewTicker(100 * time.Millisecond) // raft.go:207
	defer ticker.Stop()
	committed, term := make([]uint64, G), make([]uint64, G)
	role, out := make([]uint8, G), make([]uint32, G)
	for {
		select {
		case <-ticker.C:
			// 1. every group's inbound messages -> ONE sparse inbox post (MsgApp resolved against that group's log)
			var msgs []mrq.Msg
		drain:
			for {
				select {
				case gm := <-rc.inbox:
					msgs = append(msgs, rc.resolve(gm.group, gm.m))
				default:
					break drain
				}
			}
			// 2. proposals of every group -> ONE eng.Propose (leaders only; followers forward, see hostnode.py)
		collect:
			for {
				select {
				case p := <-props:
					rc.groups[p.g].pending = append(rc.groups[p.g].pending, []byte(p.s))
				default:
					break collect
				}
			}
			var pg []uint64
			var pn []uint32
			for g := range rc.groups {
				gs := &rc.groups[g]
				gs.posted = 0
				if gs.role == mrq.RoleLeader && len(gs.pending) > 0 {
					gs.posted = len(gs.pending)
					if gs.posted > 255 {
						gs.posted = 255
					}
					pg, pn = append(pg, uint64(g)), append(pn, uint32(gs.posted))
				}
			}
			if err := rc.eng.Step(0, msgs); err != nil {
				rc.fail(err)
				return
			}
			if len(pg) > 0 {
				rc.eng.Propose(0, pg, pn)
			}
			// 3. ONE tick for all groups
			if err := rc.eng.Tick(0); err != nil {
				rc.fail(err)
				return
			}
			// 4. Ready for all groups from one export, then per-group handling and the commit demux
			if err := rc.eng.Ready(committed, term, role, out); err != nil {
				rc.fail(err)
				return
			}
			for g := range rc.groups {
				if !rc.handleReady(g, committed[g], term[g], role[g], out[g]) {
					return
				}
			}
		case <-rc.stopc:
			rc.stop()
			return
		}
	}
}

// resolve: as (*raftNode).resolve in raftpipe_mrq.go, against group g's log, with Msg.Group = g.
func (rc *multiRaftNode) resolve(g uint64, m inbound) mrq.Msg {
	one := raftNode{log: rc.groups[g].log}
	msg := one.resolve(m)
	rc.groups[g].log = one.log
	msg.Group = g
	return msg
}

// handleReady: (*raftNode).handleReady for group g; committed[g] advances become sends on commitC[g], in
// log order, empty entries skipped (raft.go:84-86).
func (rc *multiRaftNode) handleReady(g int, committed, term uint64, role uint8, out uint32) bool {
	gs := &rc.groups[g]
	if role == mrq.RoleLeader {
		if out&mrq.OutBecameLeader != 0 {
			gs.log = append(gs.log, entry{term: term})
		}
		for _, p := range gs.pending[:gs.posted] {
			gs.log = append(gs.log, entry{term: term, data: p})
		}
		gs.pending = gs.pending[gs.posted:]
	}
	gs.term, gs.role = term, role
	// wal.Save(HardState, new entries) for this group; transport.Send(messages rebuilt from out) — hostnode.py
	for gs.applied < committed && gs.applied < uint64(len(gs.log)) {
		gs.applied++
		if d := gs.log[gs.applied-1].data; len(d) > 0 {
			s := string(d)
			select {
			case rc.commitC[g] <- &s:
			case <-rc.stopc:
				rc.stop()
				return false
			}
		}
	}
	return true
}

func (rc *multiRaftNode) fail(err error) { // writeError (raft.go:136-142), for every group
	for _, c := range rc.commitC {
		close(c)
	}
	rc.errorC <- err
	close(rc.errorC)
	if rc.eng != nil {
		rc.eng.Close()
	}
}

func (rc *multiRaftNode) stop() { // raft.go:191-196
	for _, c := range rc.commitC {
		close(c)
	}
	close(rc.errorC)
	rc.eng.Close()
}
