// GpuIndexingWorkerActor.scala -- drop-in replacement of cpslab.deploy.server.IndexingWorkerActor
// that keeps the actor's message protocol and moves the inverted index + scoring loop to the GPU
// through integration/jni/apss_jni.c -> include/apss.h.
//
// UNVERIFIED SOURCE: there is no JVM / scalac in the build image; this file has never been compiled.
// The behaviour it encodes is exercised (in Python, against the same C ABI) by
// all-pairs-similarity_b200/worker.py and tests/test_worker_mirror.py.
//
// Wiring (two edits in the reference):
//   EntryProxyActor.scala:116 and :135
//     - context.actorOf(Props(new IndexingWorkerActor(conf)))
//     + context.actorOf(Props(new GpuIndexingWorkerActor(conf)))
//   EntryProxyActor.scala:113-122 (handleDataPacket): stop splitting by dimension (spawnToIndexActor,
//   :37-49) and forward the whole DataPacket to ONE GPU worker:  gpuWorker ! IndexData(dp.vectors)
//   (an id-range shard holds all dimensions of its vectors, so every pair is scored exactly once).
package cpslab.deploy.server

import scala.collection.JavaConverters._
import scala.collection.mutable
import scala.collection.mutable.ArrayBuffer
import scala.concurrent.duration._
import scala.language.postfixOps

import akka.actor.{Actor, ActorSelection, Cancellable, ReceiveTimeout}
import com.typesafe.config.Config
import cpslab.message._
import cpslab.vector.SparseVectorWrapper

import cpslab.gpu.ApssNative     // integration/java/cpslab/gpu/ApssNative.java: static natives <-> integration/jni/apss_jni.c

private class GpuIndexingWorkerActor(conf: Config) extends Actor {
  // same keys as IndexingWorkerActor.scala:23,26,33,44 and WriteWorkerActor.scala:31,35
  val similarityThreshold = conf.getDouble("cpslab.allpair.similarityThreshold")
  val outputWritingDuration = conf.getLong("cpslab.allpair.outputIODuration")
  private val expDuration = conf.getLong("cpslab.allpair.benchmark.expDuration")
  private val vectorDim = conf.getInt("cpslab.allpair.vectorDim")
  private val indexThreshold =
    if (conf.hasPath("cpslab.allpair.indexThreshold")) conf.getDouble("cpslab.allpair.indexThreshold") else 0.0
  // the GPUs that share this worker's index (id-range shards, dispatched below the C ABI); default: GPU 0 alone
  private val devices: Array[Int] =
    if (conf.hasPath("cpslab.allpair.gpu.devices")) conf.getIntList("cpslab.allpair.gpu.devices").asScala.map(_.intValue).toArray
    else Array(0)
  // 0 = every posting visited (counters as in the reference); 3 = exact index reduction, same pairs (include/apss.h)
  private val pruning = if (conf.hasPath("cpslab.allpair.gpu.pruning")) conf.getInt("cpslab.allpair.gpu.pruning") else 0
  // "R1" = the specified result; "R0" = as built: the first posting list of every query is walked but never scored
  // (IWA:89 + IWA:106-107), so pairs whose only shared dimensions are the FIRST element of the wrapper's Set are dropped
  private val asBuilt =
    conf.hasPath("cpslab.allpair.gpu.semantics") && conf.getString("cpslab.allpair.gpu.semantics").toUpperCase == "R0"

  val writeBuffer = new mutable.HashMap[String, mutable.HashMap[String, Double]]
  var replyTo: Option[ActorSelection] = None
  var ioTask: Cancellable = null
  private var stopUpdateIndex = false
  private val ids = new ArrayBuffer[String]                       // internal id -> caller's String id
  private val firstOf = new mutable.HashMap[String, Long]         // String id -> key (first internal id)
  private var dups = false
  private val handle = ApssNative.create(vectorDim, similarityThreshold, indexThreshold, devices,
    if (asBuilt) ApssNative.SEM_R0 else ApssNative.SEM_R1, pruning)

  if (expDuration > 0) context.setReceiveTimeout(expDuration milliseconds)   // IWA:37-39

  override def preStart(): Unit = {                                          // IWA:41-51
    val system = context.system
    import system.dispatcher
    replyTo = Some(context.actorSelection(conf.getString("cpslab.allpair.outputActor")))
    if (outputWritingDuration > 0) {
      ioTask = context.system.scheduler.schedule(0 milliseconds, outputWritingDuration milliseconds, self, IOTicket)
    }
  }

  override def postStop(): Unit = ApssNative.destroy(handle)

  // buildInvertedIndex + querySimilarItems (IWA:61-111) on the GPU
  // `firsts`: per vector the first element of the wrapper's Set[Int] in ITS OWN iteration order -- the very Set the
  // reference iterates at IWA:102, so no emulation of the Scala collection order is involved here
  private def queryAndIndex(vectors: Seq[(String, cpslab.vector.SparseVector)], firsts: Array[Int], skipAdmit: Boolean):
  mutable.HashMap[String, mutable.HashMap[String, Double]] = {
    val n = vectors.size
    val indptr = new Array[Long](n + 1)
    for (i <- 0 until n) {
      require(vectors(i)._2.size == vectorDim, s"vector1 size: ${vectors(i)._2.size}, vector2 size: $vectorDim")  // CU:99
      indptr(i + 1) = indptr(i) + vectors(i)._2.indices.length
    }
    val indices = new Array[Int](indptr(n).toInt)
    val values = new Array[Double](indptr(n).toInt)
    val keys = new Array[Long](n)
    val base = ids.size
    // keys of this batch: a String id seen before keeps the key of its first occurrence (IWA:91 compares Strings).
    // Nothing is recorded in firstOf / ids until the native call has succeeded (a refused batch leaves no trace).
    val fresh = new mutable.HashMap[String, Long]
    var dupsNow = dups
    for (i <- 0 until n) {
      val v = vectors(i)._2
      System.arraycopy(v.indices, 0, indices, indptr(i).toInt, v.indices.length)
      System.arraycopy(v.values, 0, values, indptr(i).toInt, v.values.length)
      val id = vectors(i)._1
      if (firstOf.contains(id) || fresh.contains(id)) dupsNow = true
      keys(i) = firstOf.getOrElse(id, fresh.getOrElseUpdate(id, (base + i).toLong))
    }
    val flags = (if (stopUpdateIndex) ApssNative.QUERY_ONLY else 0) | (if (skipAdmit) ApssNative.SKIP_ADMIT else 0)
    val res = ApssNative.insertBatch(handle, indptr, indices, values, if (dupsNow) keys else null,
      if (asBuilt) firsts else null, flags)                     // throws RuntimeException: batch dropped, index unchanged
    val nPairs = res(1).toInt
    val status = new Array[Byte](n)
    ApssNative.fetchStatus(handle, status)
    val q = new Array[Int](nPairs); val c = new Array[Int](nPairs); val sim = new Array[Double](nPairs)
    ApssNative.fetchPairs(handle, q, c, sim)
    dups = dupsNow
    if (!stopUpdateIndex) { vectors.foreach(v => ids += v._1); firstOf ++= fresh }
    val out = new mutable.HashMap[String, mutable.HashMap[String, Double]]
    for (i <- 0 until n if status(i) == 2) out.getOrElseUpdate(vectors(i)._1, new mutable.HashMap[String, Double])  // IWA:106
    for (k <- 0 until nPairs) out(vectors(q(k))._1) += ids(c(k)) -> sim(k)
    out
  }

  private def handleBatch(vectors: Seq[(String, cpslab.vector.SparseVector)], firsts: Array[Int], skipAdmit: Boolean): Unit = {
    try {                                                                                    // IWA:124
      val out = queryAndIndex(vectors, firsts, skipAdmit)
      if (replyTo.isDefined) {
        if (outputWritingDuration <= 0) {
          replyTo.get ! SimilarityOutput(out, System.currentTimeMillis())                    // IWA:130
        } else {
          for ((qid, sims) <- out; (cid, s) <- sims)                                         // IWA:113-120
            writeBuffer.getOrElseUpdate(qid, new mutable.HashMap[String, Double]) += cid -> s
        }
      }
    } catch {
      case e: Exception => e.printStackTrace()                                               // IWA:135-137
    }
  }

  def receive: Receive = {
    case IndexData(vectors) =>                       // wrappers carry admitted, pruned vectors (EPA:97, WWA:192-194)
      val ws = vectors.toSeq
      handleBatch(ws.map(_.sparseVector), ws.map(w => if (w.indices.isEmpty) -1 else w.indices.head).toArray, skipAdmit = true)
    case IOTicket =>                                                                         // IWA:138-142
      if (!writeBuffer.isEmpty) {
        replyTo.get ! SimilarityOutput(writeBuffer.clone(), System.currentTimeMillis())
        writeBuffer.clear()
      }
    case ReceiveTimeout =>                                                                   // IWA:143-144
      stopUpdateIndex = true
      ApssNative.freeze(handle)
    case t @ Test(_) =>                                                                      // IWA:145-147
      replyTo.get ! t
  }
}
