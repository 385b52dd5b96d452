package org.apache.bifromq.dist.worker.gpumatch;

import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.nio.charset.StandardCharsets;

/** (blob, int64 offsets[n + 1]) marshalling of UTF-8 strings into direct buffers, the layout include/bfq_gpumatch.h documents. */
final class Blobs {
    final ByteBuffer blob;
    final ByteBuffer off;
    final ByteBuffer zeros;   // int32[n] of zeros: "every topic belongs to tenant 0 of the list"

    private Blobs(ByteBuffer blob, ByteBuffer off, ByteBuffer zeros) {
        this.blob = blob;
        this.off = off;
        this.zeros = zeros;
    }

    static Blobs ofUtf8(String... strings) {
        byte[][] bytes = new byte[strings.length][];
        long total = 0;
        for (int i = 0; i < strings.length; i++) {
            bytes[i] = strings[i].getBytes(StandardCharsets.UTF_8);
            total += bytes[i].length;
        }
        ByteBuffer blob = ByteBuffer.allocateDirect((int) Math.max(total, 1));
        ByteBuffer off = ByteBuffer.allocateDirect(8 * (strings.length + 1)).order(ByteOrder.LITTLE_ENDIAN);
        long at = 0;
        off.putLong(0, 0L);
        for (int i = 0; i < strings.length; i++) {
            blob.put(bytes[i]);
            at += bytes[i].length;
            off.putLong(8 * (i + 1), at);
        }
        blob.flip();
        return new Blobs(blob, off, ByteBuffer.allocateDirect(4 * Math.max(strings.length, 1)));
    }
}
