/*
 * Small helpers around MvDeconB200 for the GUI (DeconvolutionGUI.patch).
 */
package net.preibisch.mvrecon.process.cuda;

public class MvDeconB200Tools
{
	/** libmvdecon transforms every length 2^a 3^b 5^c from 32 to 1152 (include/mvdecon.h, mvd_supported_fft_lengths) */
	public static boolean isSupportedLength( int n )
	{
		if ( n < 32 || n > 1152 )
			return false;

		for ( final int p : new int[]{ 2, 3, 5 } )
			while ( n % p == 0 )
				n /= p;

		return n == 1;
	}

	public static boolean isSupportedBlock( final int[] blockSize )
	{
		for ( final int b : blockSize )
			if ( !isSupportedLength( b ) )
				return false;

		return true;
	}
}
