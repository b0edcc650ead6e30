/*
 * Factory of the level-2 block threads; mirrors ComputeBlockSeqThreadCUDAFactory.java:33-79 (one thread per selected device).
 */
package net.preibisch.mvrecon.process.deconvolution.iteration.sequential;

import java.util.HashMap;

import net.preibisch.mvrecon.process.cuda.CUDADevice;
import net.preibisch.mvrecon.process.cuda.MvDeconB200;
import net.preibisch.mvrecon.process.deconvolution.iteration.ComputeBlockThreadFactory;

public class ComputeBlockSeqThreadB200Factory implements ComputeBlockThreadFactory< ComputeBlockSeqThread >
{
	final float minValue;
	final float lambda;
	final int[] blockSize;
	final MvDeconB200 lib;
	final HashMap< Integer, CUDADevice > idToCudaDevice;

	public ComputeBlockSeqThreadB200Factory( final float minValue, final float lambda, final int[] blockSize, final MvDeconB200 lib,
			final HashMap< Integer, CUDADevice > idToCudaDevice )
	{
		this.minValue = minValue;
		this.lambda = lambda;
		this.blockSize = blockSize.clone();
		this.lib = lib;
		this.idToCudaDevice = idToCudaDevice;
	}

	@Override
	public ComputeBlockSeqThread create( final int id )
	{
		return new ComputeBlockSeqThreadB200( minValue, lambda, id, blockSize, lib, idToCudaDevice.get( id ) );
	}

	@Override
	public int numParallelBlocks() { return idToCudaDevice.keySet().size(); }

	/** the loaded library and the first selected device, for the resident driver (MultiViewDeconvolutionB200) */
	public MvDeconB200 lib() { return lib; }
	public int firstDeviceId() { return idToCudaDevice.get( 0 ).getDeviceId(); }

	@Override
	public String toString()
	{
		String out = "B200 native (fused passes, libmvdecon " + lib.mvd_version() + ") using " + numParallelBlocks() + " devices:";

		for ( int devId = 0; devId < numParallelBlocks(); ++devId )
			out += " [" + idToCudaDevice.get( devId ) + "]";

		return out;
	}
}
