! propack_b200.f90 -- ISO_C_BINDING interfaces to the non-Fortran-ABI part of libpropack_b200.so.
!
! The drivers themselves (DLANSVD, DLANSVD_IRL, DLANBPRO, DREORTH, DGETU0, ... and the S/C/Z variants) are exported with
! the gfortran Fortran-77 ABI under the reference's own names (double/dlansvd.F:1-3 etc.), so existing PROPACK callers
! need no interface at all: they just link this library instead of libdpropack_<plat>.a.  This module only adds what a
! caller needs to keep the matrix on the GPU: register it once, pass PROPACK_B200_APROD_D as the APROD argument and the
! handle in IPARM(1).
!
! NOTE: there is no Fortran compiler in the build image, so this file has never been compiled; it is the binding a
! maintainer would add (INTEGRATION.md section 2).  Signatures mirror include/propack_b200.h.
module propack_b200
  use iso_c_binding
  implicit none
  interface
    ! int propack_b200_init(void)
    integer(c_int) function propack_b200_init() bind(C, name='propack_b200_init')
      import :: c_int
    end function
    ! int propack_b200_csr_create_{s,d,c,z}(int m, int n, const int* rowptr, const int* colind, const T* values, int base)
    integer(c_int) function propack_b200_csr_create_s(m, n, rowptr, colind, values, index_base) &
        bind(C, name='propack_b200_csr_create_s')
      import :: c_int, c_float
      integer(c_int), value :: m, n, index_base
      integer(c_int), intent(in) :: rowptr(*), colind(*)
      real(c_float), intent(in) :: values(*)
    end function
    integer(c_int) function propack_b200_csr_create_d(m, n, rowptr, colind, values, index_base) &
        bind(C, name='propack_b200_csr_create_d')
      import :: c_int, c_double
      integer(c_int), value :: m, n, index_base
      integer(c_int), intent(in) :: rowptr(*), colind(*)
      real(c_double), intent(in) :: values(*)
    end function
    integer(c_int) function propack_b200_csr_create_c(m, n, rowptr, colind, values, index_base) &
        bind(C, name='propack_b200_csr_create_c')
      import :: c_int, c_float_complex
      integer(c_int), value :: m, n, index_base
      integer(c_int), intent(in) :: rowptr(*), colind(*)
      complex(c_float_complex), intent(in) :: values(*)
    end function
    integer(c_int) function propack_b200_csr_create_z(m, n, rowptr, colind, values, index_base) &
        bind(C, name='propack_b200_csr_create_z')
      import :: c_int, c_double_complex
      integer(c_int), value :: m, n, index_base
      integer(c_int), intent(in) :: rowptr(*), colind(*)
      complex(c_double_complex), intent(in) :: values(*)
    end function
    ! int propack_b200_dense_create_d(int m, int n, const double* A, long lda)   (column-major, as Fortran stores it)
    integer(c_int) function propack_b200_dense_create_d(m, n, a, lda) bind(C, name='propack_b200_dense_create_d')
      import :: c_int, c_long, c_double
      integer(c_int), value :: m, n
      integer(c_long), value :: lda
      real(c_double), intent(in) :: a(lda, *)
    end function
    ! ---- multi-GPU, one process (one MPI rank) per GPU: include/propack_b200.h "multi-GPU" ----------------------------
    ! int propack_b200_comm_unique_id(void* id128_out)            rank 0 only; broadcast the 128 bytes (e.g. MPI_Bcast)
    integer(c_int) function propack_b200_comm_unique_id(id128) bind(C, name='propack_b200_comm_unique_id')
      import :: c_int, c_char
      character(kind=c_char), intent(out) :: id128(128)
    end function
    ! int propack_b200_comm_init(int rank, int world, const void* id128)        collective
    integer(c_int) function propack_b200_comm_init(rank, world, id128) bind(C, name='propack_b200_comm_init')
      import :: c_int, c_char
      integer(c_int), value :: rank, world
      character(kind=c_char), intent(in) :: id128(128)
    end function
    integer(c_int) function propack_b200_comm_finalize() bind(C, name='propack_b200_comm_finalize')
      import :: c_int
    end function
    ! void propack_b200_shard_bounds(long dim, int world, int rank, long* lo, long* hi)     rank owns [lo, hi), 0-based
    subroutine propack_b200_shard_bounds(dim, world, rank, lo, hi) bind(C, name='propack_b200_shard_bounds')
      import :: c_int, c_long
      integer(c_long), value :: dim
      integer(c_int), value :: world, rank
      integer(c_long), intent(out) :: lo, hi
    end subroutine
    ! int propack_b200_csr_create_sharded_d(int mg, int ng, row_rowptr, row_colind, row_values, colt_rowptr, colt_rowind,
    !                                       colt_values, int index_base)      this rank's rows of A and columns of A (transposed)
    integer(c_int) function propack_b200_csr_create_sharded_d(mg, ng, row_rowptr, row_colind, row_values, colt_rowptr, &
        colt_rowind, colt_values, index_base) bind(C, name='propack_b200_csr_create_sharded_d')
      import :: c_int, c_double
      integer(c_int), value :: mg, ng, index_base
      integer(c_int), intent(in) :: row_rowptr(*), row_colind(*), colt_rowptr(*), colt_rowind(*)
      real(c_double), intent(in) :: row_values(*), colt_values(*)
    end function
    ! int propack_b200_dense_create_sharded_d(int mg, int ng, const double* A_rows, long lda)      this rank's rows of a dense A
    integer(c_int) function propack_b200_dense_create_sharded_d(mg, ng, a_rows, lda) &
        bind(C, name='propack_b200_dense_create_sharded_d')
      import :: c_int, c_long, c_double
      integer(c_int), value :: mg, ng
      integer(c_long), value :: lda
      real(c_double), intent(in) :: a_rows(lda, *)
    end function
    ! int propack_b200_op_destroy(int handle)
    integer(c_int) function propack_b200_op_destroy(handle) bind(C, name='propack_b200_op_destroy')
      import :: c_int
      integer(c_int), value :: handle
    end function
    ! const char* propack_b200_last_error(void)
    type(c_ptr) function propack_b200_last_error() bind(C, name='propack_b200_last_error')
      import :: c_ptr
    end function
  end interface
  ! The built-in APROD callbacks are ordinary Fortran-ABI externals (trailing underscore added by the compiler):
  !   external propack_b200_aprod_s, propack_b200_aprod_d, propack_b200_aprod_c, propack_b200_aprod_z
  ! e.g.   call dlansvd('y','y',m,n,k,kmax,propack_b200_aprod_d,U,ldu,Sigma,bnd,V,ldv,tolin, &
  !                     work,lwork,iwork,liwork,doption,ioption,info,dparm,iparm)      with iparm(1) = handle
end module propack_b200
